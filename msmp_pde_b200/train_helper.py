"""``training_loop`` of the reference (experiments/train_helper.py:66-148) for the drop-in GNN solvers.

Same signature, same order of ``random`` draws (number of unrollings, then the start indices), same pushforward trick
(``:106-112``: the unrolled predictions are computed without gradients and fed back through
``create_next_graph``), same loss ``sqrt(criterion(pred, graph.y))`` and the same returned ``losses / batch_size``.

What differs is the execution of the gradient step: with the optimizer the reference's scripts build
(``optim.AdamW(model.parameters(), lr=args.lr)``, experiments/train.py:410 -- or any capturable optimizer) and graphs
that keep one topology (fixed grid and batch size, which is what ``GraphCreator`` produces), forward + loss + backward
+ optimizer update run as ONE CUDA-graph replay (``GraphedTrainStep``); the graph fields are copied into its static
buffers and the optimizer's hyper-parameters are re-read every step, so ``MultiStepLR`` (train.py:411,437) acts as in the
reference.  Otherwise the step runs eagerly -- still on the CUDA kernels -- exactly like the reference loop.
"""
from __future__ import annotations

import random

import torch

from . import optim as moptim
from .models_gnn import MP_PDE_SolverLEMLinGatedSave
from .train_step import GraphedForward, GraphedTrainStep


def reset_state_bool(model) -> bool:
    """experiments/train_helper.py:10-13"""
    inst = isinstance(model, MP_PDE_SolverLEMLinGatedSave)
    has = hasattr(model, "save_state")
    return inst or (has and model.save_state is not None and model.save_state)


def _is_sum_mse(criterion) -> bool:
    return isinstance(criterion, torch.nn.MSELoss) and criterion.reduction == "sum"


def training_loop(model: torch.nn.Module, unrolling: list, batch_size: int, optimizer, loader, graph_creator, criterion,
                  device="cpu") -> torch.Tensor:
    """One training epoch with random starting points for every trajectory (experiments/train_helper.py:66-148)."""
    if f"{model}" != "GNN":
        raise NotImplementedError("msmp_pde_b200.train_helper.training_loop drives the GNN solvers only")
    capturable = bool(optimizer.param_groups and all(g.get("capturable", False) for g in optimizer.param_groups))
    fused_step = torch.device(device).type == "cuda" and _is_sum_mse(criterion) and not reset_state_bool(model) \
        and (capturable or moptim.supported(optimizer))
    cache = model.__dict__.setdefault("_msmp_graphed_steps", {})
    losses = []
    for (u_base, u_super, x, variables) in loader:
        if not fused_step:
            optimizer.zero_grad()          # (the captured step overwrites every gradient; its buffers must stay allocated)
        # graphs are built on the target device: the creator then hands out the same device-resident edge list for
        # every batch and the model's topology cache / the captured step are reused.  The grid coordinates x stay on
        # the host: the creator keys its topology cache on their bytes (no device round trip) and moves one row itself.
        u_super = u_super.to(device)
        # Randomly choose number of unrollings, then the starting (time) points on the solution manifold
        unrolled_graphs = random.choice(unrolling)
        steps = [t for t in range(graph_creator.tw,
                                  graph_creator.t_res - graph_creator.tw - (graph_creator.tw * unrolled_graphs) + 1)]
        random_steps = random.choices(steps, k=batch_size)
        data, labels = graph_creator.create_data(u_super, random_steps)
        graph = graph_creator.create_graph(data, labels, x, variables, random_steps).to(device)

        with torch.no_grad():                               # the pushforward trick
            for _ in range(unrolled_graphs):
                random_steps = [rs + graph_creator.tw for rs in random_steps]
                _, labels = graph_creator.create_data(u_super, random_steps)
                pred = model(graph)
                graph = graph_creator.create_next_graph(graph, pred, labels, random_steps).to(device)

        if fused_step:
            # keyed on the storage and size of the topology (Python ids are recycled); the step itself verifies that
            # a graph arriving under the same key really has the captured edge list (GraphedTrainStep._check_topology)
            ei = graph.edge_index
            key = (id(optimizer), ei.data_ptr(), int(ei.shape[1]), tuple(graph.x.shape), str(ei.device))
            step = cache.get(key)
            if step is not None and step.opt is not optimizer:
                step = None
            if step is None:
                # the constructor's warm-up steps run on a snapshot of parameters and optimizer state, so that this
                # call still performs exactly one optimizer update
                step = GraphedTrainStep(model, optimizer, graph, warmup=1, preserve_state=True)
                cache[key] = step
            loss = step(graph).to(graph.x.dtype)
            losses.append(loss.detach().clone() / batch_size)
        else:
            pred = model(graph)
            loss = torch.sqrt(criterion(pred, graph.y))
            loss.backward()
            losses.append(loss.detach() / batch_size)
            optimizer.step()

        if reset_state_bool(model):                          # reset the hidden state of LEM for new data
            model.embedding_lem.reset_states()
    return torch.stack(losses)


# Evaluation rollouts replay the forward pass as a CUDA graph (one replay per window); "0" keeps eager launches.
EVAL_GRAPH = True


def _forward(model, graph):
    """``model(graph)`` under no_grad -- as a CUDA-graph replay when the graph lives on a CUDA device and the model keeps no
    state between calls (the stateful LEMS variants run eagerly).  One captured forward per topology is cached on the
    model; a graph with another topology or other field shapes gets its own."""
    if not (EVAL_GRAPH and graph.x.is_cuda and not reset_state_bool(model) and not torch.is_grad_enabled()):
        return model(graph)
    cache = model.__dict__.setdefault("_msmp_graphed_forwards", [])
    for gf in cache:
        if gf.matches(graph):
            return gf(graph)
    if len(cache) >= 4:
        cache.pop(0)
    gf = GraphedForward(model, graph)
    cache.append(gf)
    return gf(graph)


def test_timestep_losses(model, steps: list, batch_size: int, loader, graph_creator, criterion, device="cpu") -> None:
    """Loss of one forward pass at the time points that are multiples of the time window
    (experiments/train_helper.py:150-203); prints like the reference, returns None."""
    if f"{model}" != "GNN":
        raise NotImplementedError("msmp_pde_b200.train_helper drives the GNN solvers only")
    for step in steps:
        if step != graph_creator.tw and step % graph_creator.tw != 0:
            continue
        losses = []
        for (u_base, u_super, x, variables) in loader:
            with torch.no_grad():
                u_super = u_super.to(device)
                same_steps = [step] * batch_size
                data, labels = graph_creator.create_data(u_super, same_steps)
                graph = graph_creator.create_graph(data, labels, x, variables, same_steps).to(device)
                pred = _forward(model, graph)
                losses.append(criterion(pred, graph.y) / batch_size)
            if reset_state_bool(model):
                model.embedding_lem.reset_states()
        losses = torch.stack(losses)
        print(f'Step {step}, mean loss {torch.mean(losses)}')


def test_unrolled_losses(model, steps: list, batch_size: int, nr_gt_steps: int, nx_base_resolution: int, loader,
                         graph_creator, criterion, device="cpu") -> torch.Tensor:
    """Loss of the full autoregressive rollout of every trajectory (experiments/train_helper.py:205-292): the
    prediction of one window is the input of the next through ``create_next_graph``."""
    if f"{model}" != "GNN":
        raise NotImplementedError("msmp_pde_b200.train_helper drives the GNN solvers only")
    losses, losses_base = [], []
    tw = graph_creator.tw
    for (u_base, u_super, x, variables) in loader:
        losses_tmp, losses_base_tmp = [], []
        with torch.no_grad():
            u_base, u_super = u_base.to(device), u_super.to(device)
            same_steps = [tw * nr_gt_steps] * batch_size
            data, labels = graph_creator.create_data(u_super, same_steps)
            graph = graph_creator.create_graph(data, labels, x, variables, same_steps).to(device)
            pred = _forward(model, graph)
            losses_tmp.append(criterion(pred, graph.y) / nx_base_resolution / batch_size)
            # unroll the trajectory; every window adds its loss
            for step in range(tw * (nr_gt_steps + 1), graph_creator.t_res - tw + 1, tw):
                same_steps = [step] * batch_size
                _, labels = graph_creator.create_data(u_super, same_steps)
                graph = graph_creator.create_next_graph(graph, pred, labels, same_steps).to(device)
                pred = _forward(model, graph)
                losses_tmp.append(criterion(pred, graph.y) / nx_base_resolution / batch_size)
            if reset_state_bool(model):
                model.embedding_lem.reset_states()
            # losses of the numerical baseline
            for step in range(tw * nr_gt_steps, graph_creator.t_res - tw + 1, tw):
                same_steps = [step] * batch_size
                _, labels_super = graph_creator.create_data(u_super, same_steps)
                _, labels_base = graph_creator.create_data(u_base, same_steps)
                losses_base_tmp.append(criterion(labels_super, labels_base) / nx_base_resolution / batch_size)
        losses.append(torch.sum(torch.stack(losses_tmp)))
        losses_base.append(torch.sum(torch.stack(losses_base_tmp)))
    losses = torch.stack(losses)
    losses_base = torch.stack(losses_base)
    print(f'Unrolled forward losses {torch.mean(losses)}')
    print(f'Unrolled forward base losses {torch.mean(losses_base)}')
    return losses


def _sq_norm_over_fields(t: torch.Tensor) -> torch.Tensor:
    if t.dim() != 4:
        raise ValueError("expected [B, n_t, d, n_x]")
    return t.sum(dim=2)          # |.|^2 on R^d  ->  [B, n_t, n_x]


def compute_spacetime_L2_norms(losses: torch.Tensor, norms: torch.Tensor):
    """Absolute and relative error in L2(Omega x [0, T]) (experiments/train_helper.py:299-329).
    ``losses = (pred - true)^2`` and ``norms = true^2``, both [B, n_t, d, n_x]; returns two scalars: the sample mean of
    sqrt(mean over space and time of the squared R^d norm), and its ratio to the same functional of ``norms``."""
    assert losses.shape == norms.shape, "loss and norms do not have the same shape"
    err = _sq_norm_over_fields(losses).mean(dim=(1, 2)).sqrt().mean()
    ref = _sq_norm_over_fields(norms).mean(dim=(1, 2)).sqrt().mean()
    return err, err / ref


def compute_space_L2_norms(losses: torch.Tensor, norms: torch.Tensor):
    """Per time point: absolute and relative error in L2(Omega) (experiments/train_helper.py:331-360); returns two
    [n_t] vectors (sample means)."""
    assert losses.shape == norms.shape, "loss and norms do not have the same shape"
    err = _sq_norm_over_fields(losses).mean(dim=2).sqrt().mean(dim=0)
    ref = _sq_norm_over_fields(norms).mean(dim=2).sqrt().mean(dim=0)
    return err, err / ref


def compute_L2_norms(model, batch_size: int, nr_gt_steps: int, loader, graph_creator, device="cpu"):
    """The metric the reference reports (experiments/train_helper.py:362-471): every trajectory is rolled out
    autoregressively over its whole length (the prediction of one window is the next window's input), the squared
    errors and squared targets of all windows are laid out as [B, n_t, d, n_x] and reduced with
    ``compute_spacetime_L2_norms``.  Prints like the reference and returns ``(L2 error, relative L2 error)`` as floats."""
    if f"{model}" != "GNN":
        raise NotImplementedError("msmp_pde_b200.train_helper drives the GNN solvers only")
    tw = graph_creator.tw
    err_all, ref_all = [], []
    for (u_base, u_super, x, variables) in loader:
        bs = u_super.size(0)
        d = u_super.size(2) if u_super.dim() == 4 else 1

        def fields(t):          # [B * n_x, d * tw] -> [B, tw, d, n_x]
            return t.reshape(bs, -1, d, tw).permute(0, 3, 2, 1)

        err_w, ref_w = [], []
        with torch.no_grad():
            u_super = u_super.to(device)
            steps = [tw * nr_gt_steps] * bs
            data, labels = graph_creator.create_data(u_super, steps)
            graph = graph_creator.create_graph(data, labels, x, variables, steps).to(device)
            pred = _forward(model, graph)
            err_w.append(fields(torch.square(pred - graph.y)))
            ref_w.append(fields(torch.square(graph.y)))
            for step in range(tw * (nr_gt_steps + 1), graph_creator.t_res - tw + 1, tw):
                steps = [step] * bs
                _, labels = graph_creator.create_data(u_super, steps)
                graph = graph_creator.create_next_graph(graph, pred, labels, steps).to(device)
                pred = _forward(model, graph)
                err_w.append(fields(torch.square(pred - graph.y)))
                ref_w.append(fields(torch.square(graph.y)))
            if reset_state_bool(model):
                model.embedding_lem.reset_states()
        err_all.append(torch.cat(err_w, 1))
        ref_all.append(torch.cat(ref_w, 1))
    l2, l2_rel = compute_spacetime_L2_norms(torch.cat(err_all, 0), torch.cat(ref_all, 0))
    print(f'L2 error {l2.item()}')
    print(f'L2 relative error {100 * l2_rel.item()} %')
    return l2.item(), l2_rel.item()


test_timestep_losses.__test__ = False      # (named as in the reference; not pytest tests)
test_unrolled_losses.__test__ = False
