"""Round-2 parity tests (B200): the benchmarked shapes at full size, the other time windows, the lem_cuda boundary, the
captured step with the optimizer the reference's train.py builds, and the NCCL data-parallel step."""
import copy
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

from tests import golden_io  # noqa: E402
from tests.test_models_gpu import GRAD_TOL, OUT_TOL, _grad_errs  # noqa: E402
from tests.test_oracle_golden import TW_CASES  # noqa: E402
from tests.util import formula_weights_, rel_err  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _model_vs_oracle(pde, data, meta, tw=25):
    from msmp_pde_b200 import models_gnn2D
    from oracle import models as om
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = models_gnn2D.MP_PDE_Solver2DLEMLinGated(pde, tw, 128, 6, meta["eq_variables"]).to(dev)
    torch.set_default_dtype(torch.float64)
    torch.set_num_threads(os.cpu_count())
    ref = om.MP_PDE_Solver2DLEMLinGated(pde, tw, 128, 6, meta["eq_variables"])
    ref.load_state_dict({k: v.double().cpu() for k, v in model.state_dict().items()})
    dd = data.clone().to(dev)
    out = model(dd)
    torch.sqrt(((out - dd.y) ** 2).sum()).backward()
    outr = ref(data)
    torch.sqrt(((outr - data.y) ** 2).sum()).backward()
    assert rel_err(out, outr) < OUT_TOL
    errs = _grad_errs(model, ref)
    worst = max(errs, key=errs.get)
    assert errs[worst] <= 1.0, (worst, errs[worst])


def test_c2_full_bench_shape():
    """BASELINE config 2 at the size bench.py runs as a sub-record: 64 graphs x 100 nodes (N = 6400, E = 37632), outputs and
    every parameter gradient against the float64 oracle."""
    from msmp_pde_b200 import synth
    _model_vs_oracle(*synth.config_c2(B=64, nx=100, seed=0))


def test_c4_full_lattice_graph():
    """One full 128 x 128 lattice graph of BASELINE config 4 (16384 nodes, 65024 edges, 128 InstanceNorm chunks per graph,
    256 LEM node tiles): outputs and gradients against the float64 oracle."""
    from msmp_pde_b200 import synth
    _model_vs_oracle(*synth.config_c4(B=1, side=128, seed=3))


@pytest.mark.parametrize("cls,fname,pde_name,eq,tw", TW_CASES)
def test_time_windows_vs_golden_and_oracle(cls, fname, pde_name, eq, tw):
    """time_window 20 / 50: generic decoder geometry (128 -> 29 -> 20, 128 -> 59 -> 50), F_u = 20 / 50 / 100 (zero padded to
    32 / 64 / 128 columns), LEM over 50 steps -- against fixtures written by the reference classes and the oracle's
    gradients."""
    import importlib
    from oracle import models as om
    dev = torch.device("cuda:0")
    g = golden_io.load(fname)
    pde, data = golden_io.model_inputs(g, pde_name)
    mod = importlib.import_module("msmp_pde_b200." + ("models_gnn2D" if "2D" in cls else "models_gnn"))
    torch.set_default_dtype(torch.float64)
    model = getattr(mod, cls)(pde, time_window=tw, hidden_features=128, hidden_layer=6, eq_variables=eq)
    formula_weights_(model)
    model = model.to(dev)
    dd = copy.copy(data).clone().to(dev)
    out = model(dd)
    loss = torch.sqrt(torch.nn.functional.mse_loss(out, dd.y, reduction="sum"))
    loss.backward()
    assert out.shape == dd.y.shape
    assert rel_err(out, torch.from_numpy(g["out"])) < OUT_TOL
    assert abs(float(loss) - float(g["loss"])) < 1e-5 * float(g["loss"])
    ref = getattr(om, cls)(pde, time_window=tw, hidden_features=128, hidden_layer=6, eq_variables=eq)
    formula_weights_(ref)
    outr = ref(data)
    torch.sqrt(torch.nn.functional.mse_loss(outr, data.y, reduction="sum")).backward()
    errs = _grad_errs(model, ref)
    worst = max(errs, key=errs.get)
    assert errs[worst] <= 1.0, (worst, errs[worst])


# ---------------------------------------------------------------------------------------------- lem_cuda boundary
class _RefLEMFunction(torch.autograd.Function):
    """The reference's LEMFunction (experiments/models_gnn.py:285-302), restated verbatim in behaviour: forwards 8
    arguments to ``lem_cuda.forward``, saves 11 tensors, hands 13 to ``lem_cuda.backward`` and returns 7 of its outputs."""
    lem_cuda = None

    @staticmethod
    def forward(ctx, inputs, weights, weights_lin_z, bias, bias_lin_z, initial_y_state, initial_z_state, dt):
        all_y, all_z, all_X, all_X2, all_ms, all_lin = _RefLEMFunction.lem_cuda.forward(
            inputs, weights, weights_lin_z, bias, bias_lin_z, initial_y_state, initial_z_state, dt)
        ctx.save_for_backward(all_X, all_X2, all_ms, all_lin, weights, weights_lin_z, bias, bias_lin_z,
                              initial_y_state, initial_z_state, dt)
        return all_y, all_z

    @staticmethod
    def backward(ctx, grad_y_states, grad_z_states):
        outputs = _RefLEMFunction.lem_cuda.backward(grad_y_states.contiguous(), grad_z_states.contiguous(),
                                                    *ctx.saved_tensors)
        d_inputs, d_w, d_wz, d_b, d_bz, d_y0, d_z0 = outputs
        return None, d_w, d_wz, d_b, d_bz, d_y0, d_z0, None


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_lem_cuda_module_contract(dtype):
    """``msmp_pde_b200.compat.lem_cuda`` behind the reference's own call pattern: 8 in / 6 out, 13 in / 7 out, all T
    states returned, gradients wrt weights, biases and the initial state against the float64 oracle recurrence; float64
    tensors (the reference's default dtype) are accepted and returned."""
    import msmp_pde_b200
    from msmp_pde_b200.compat import lem_cuda
    from oracle.models import lem_forward
    sys.modules.pop("lem_cuda", None)
    msmp_pde_b200.install()
    assert sys.modules["lem_cuda"] is lem_cuda
    _RefLEMFunction.lem_cuda = lem_cuda
    dev = torch.device("cuda:0")
    T, N, ninp, Hh = 25, 300, 6, 128
    g = torch.Generator().manual_seed(3)
    mk = lambda *s: ((torch.rand(*s, generator=g, dtype=torch.float64) * 2 - 1) / Hh ** 0.5)
    cpu = dict(inputs=torch.randn(T, N, ninp, generator=g, dtype=torch.float64), W=mk(3 * Hh, ninp + Hh),
               Wz=mk(Hh, ninp + Hh), b=mk(3 * Hh), bz=mk(Hh), y0=0.3 * torch.randn(N, Hh, generator=g, dtype=torch.float64),
               z0=0.3 * torch.randn(N, Hh, generator=g, dtype=torch.float64))
    wy, wz = torch.randn(T, N, Hh, generator=g, dtype=torch.float64), torch.randn(T, N, Hh, generator=g, dtype=torch.float64)
    # oracle
    leaves = {k: v.clone().requires_grad_(k != "inputs") for k, v in cpu.items()}
    ys, zs = lem_forward(leaves["inputs"], leaves["W"], leaves["Wz"], leaves["b"], leaves["bz"], leaves["y0"], leaves["z0"], 1.0)
    ((ys * wy).sum() + (zs * wz).sum()).backward()
    # the shim through the reference's Function
    d = {k: v.to(dev, dtype).requires_grad_(k != "inputs") for k, v in cpu.items()}
    dt = torch.tensor(1.0, dtype=dtype).reshape(1, 1).to(dev)
    all_y, all_z = _RefLEMFunction.apply(d["inputs"], d["W"], d["Wz"], d["b"], d["bz"], d["y0"], d["z0"], dt)
    assert all_y.shape == (T, N, Hh) and all_y.dtype == dtype
    ((all_y * wy.to(dev, dtype)).sum() + (all_z * wz.to(dev, dtype)).sum()).backward()
    assert rel_err(all_y, ys) < OUT_TOL and rel_err(all_z, zs) < OUT_TOL
    gscale = max(float(leaves[k].grad.abs().max()) for k in ("W", "Wz", "b", "bz", "y0", "z0"))
    for k in ("W", "Wz", "b", "bz", "y0", "z0"):
        ref = leaves[k].grad
        allow = GRAD_TOL * float(ref.abs().max()) + 1e-6 * gscale
        assert d[k].grad.dtype == dtype
        assert float((d[k].grad.double().cpu() - ref).abs().max()) <= allow, k


# ------------------------------------------------------------------------- captured step with train.py's optimizer
def test_captured_step_with_plain_adamw_and_scheduler():
    """experiments/train.py:410-411 builds ``optim.AdamW(model.parameters(), lr)`` (not capturable) and a MultiStepLR
    that rewrites param_group['lr'].  The captured step must (a) accept that optimizer, (b) follow the scheduler, and
    (c) leave the optimizer's own state where torch's eager step would have: three epochs of two steps each against the
    eager reference loop on a copy of the model, then one more eager ``optimizer.step()`` on both."""
    from msmp_pde_b200 import models_gnn2D, synth
    from msmp_pde_b200.train_step import GraphedTrainStep
    dev = torch.device("cuda:0")
    pde, data, meta = synth.config_c2(B=3, nx=50, seed=9)
    torch.manual_seed(2)
    model = models_gnn2D.MP_PDE_Solver2DLEMLinGated(pde, 25, 128, 6, meta["eq_variables"]).to(dev)
    ref = copy.deepcopy(model)
    g = data.clone().to(dev)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
    sched = torch.optim.lr_scheduler.MultiStepLR(opt, milestones=[1, 2], gamma=0.4)
    opt_r = torch.optim.AdamW(ref.parameters(), lr=1e-3)
    sched_r = torch.optim.lr_scheduler.MultiStepLR(opt_r, milestones=[1, 2], gamma=0.4)
    step = GraphedTrainStep(model, opt, g, warmup=1, preserve_state=True)
    assert step.fused is not None and step.graph is not None
    crit = torch.nn.MSELoss(reduction="sum")
    for epoch in range(3):
        for _ in range(2):
            loss = step(g)
            opt_r.zero_grad()
            loss_r = torch.sqrt(crit(ref(g), g.y))
            loss_r.backward()
            opt_r.step()
            assert abs(float(loss) - float(loss_r)) < 2e-5 * float(loss_r)
        sched.step()
        sched_r.step()
    assert opt.param_groups[0]["lr"] == pytest.approx(1e-3 * 0.16)
    sd = opt.state_dict()                                             # flushes the step counters
    assert all(float(s["step"]) == 6.0 for s in sd["state"].values())
    # Adam's update lr * m / (sqrt(v) + eps) amplifies the fp32 noise of entries whose gradient is numerically zero (up
    # to 2 lr per step if the sign flips), so the trajectories are compared in units of the summed learning rate; the
    # update kernel itself is held to torch's on identical gradients in test_fused_adamw_equals_torch_adamw.
    lr_sum = 2 * 1e-3 * (1 + 0.4 + 0.16)
    for (n, p), (_, q) in zip(model.named_parameters(), ref.named_parameters()):
        assert float((p - q).abs().max()) < 0.1 * lr_sum, n
    # hand over to torch's own eager step: same state => same next update
    for m, o in ((model, opt), (ref, opt_r)):
        o.zero_grad(set_to_none=False)
        torch.sqrt(crit(m(g), g.y)).backward()
        o.step()
    for (n, p), (_, q) in zip(model.named_parameters(), ref.named_parameters()):
        assert float((p - q).abs().max()) < 0.1 * (lr_sum + 1e-3 * 0.16), n


def test_fused_adamw_equals_torch_adamw():
    """msmp_adamw_run against torch.optim.AdamW on IDENTICAL gradients: two param groups with different lr / weight decay /
    betas, a scheduler step in between, a gradient scale; parameters and both moments after 5 steps, then a hand-over
    to torch's own step()."""
    from msmp_pde_b200.optim import FusedAdamW
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(5)
    shapes = [(128, 310), (128,), (7, 3, 5), (1,), (1025,)]
    mk = lambda: [torch.nn.Parameter(torch.randn(*s, device=dev, generator=g)) for s in shapes]
    pa = mk()
    pb = [torch.nn.Parameter(p.detach().clone()) for p in pa]
    groups = lambda ps: [dict(params=ps[:3], lr=1e-2, weight_decay=0.1, betas=(0.8, 0.95)), dict(params=ps[3:], lr=3e-3)]
    oa, ob = torch.optim.AdamW(groups(pa)), torch.optim.AdamW(groups(pb))
    for p in pa:
        p.grad = torch.zeros_like(p)
    fused = FusedAdamW(oa)
    scale = torch.tensor(0.37, device=dev)
    for it in range(5):
        for p, q in zip(pa, pb):
            gr = torch.randn(p.shape, device=dev, generator=g) * (10.0 ** (it - 2))
            p.grad.copy_(gr)
            q.grad = gr * 0.37
        fused.host_update()
        fused.launch(scale)
        ob.step()
        if it == 2:
            for o in (oa, ob):
                for grp in o.param_groups:
                    grp["lr"] *= 0.5
    for p, q in zip(pa, pb):
        assert rel_err(p, q) < 2e-6
        assert rel_err(p.grad, q.grad) < 1e-6                     # the scaled gradient is written back
        assert rel_err(oa.state[p]["exp_avg"], ob.state[q]["exp_avg"]) < 2e-6
        assert rel_err(oa.state[p]["exp_avg_sq"], ob.state[q]["exp_avg_sq"]) < 2e-6
    for p, q in zip(pa, pb):
        p.grad.fill_(0.01)
        q.grad.fill_(0.01)
    oa.step()                                                        # pre-hook flushes the step counters
    ob.step()
    assert all(float(oa.state[p]["step"]) == 6.0 for p in pa)
    for p, q in zip(pa, pb):
        assert rel_err(p, q) < 2e-6


def test_captured_step_rejects_other_topology():
    from msmp_pde_b200 import models_gnn, synth
    from msmp_pde_b200.train_step import GraphedTrainStep
    dev = torch.device("cuda:0")
    pde, data, meta = synth.config_c1(B=2, nx=40, seed=1)
    torch.manual_seed(0)
    model = models_gnn.MP_PDE_Solver(pde, 25, 128, 6, {}).to(dev)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4)
    g = data.clone().to(dev)
    step = GraphedTrainStep(model, opt, g, warmup=1)
    same = g.clone()                                  # equal edge list in another tensor: accepted (checked once)
    step(same)
    _, other, _ = synth.config_c1(B=2, nx=40, seed=1, neighbors=2)      # same node count, different edge list
    with pytest.raises(ValueError):
        step(other.clone().to(dev))


# ------------------------------------------------------------------------------------------------- NCCL data parallel
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (run with gpurun --gpus 2)")
def test_dp_nccl_step_equals_single_gpu():
    """SURVEY section 4 'Multi-GPU': the 3-graphs-per-rank NCCL step (sharded graphs, one all-reduce of the gradient bucket
    that also carries the squared error) against the single-GPU full-batch step: loss, gradient and updated weights."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tests", "dp_nccl_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("DPCHECK ")][-1]
    res = json.loads(line[len("DPCHECK "):])
    for mode, r in res.items():
        assert r["loss_rel_err"] < 1e-6, (mode, r)
        assert r["grad_max_rel_err"] < 1e-5, (mode, r)
        assert r["weight_max_abs_err_over_lr"] <= 2.0 + 1e-6 and r["weight_mismatch_frac"] < 1e-3, (mode, r)


# ------------------------------------------------------------------------------------------- reduced-precision mode
BF16_OUT_TOL = 5e-4          # of max|ref|  (measured 6e-5 .. 7e-5 on these fixtures)
BF16_GRAD_TOL = 3e-2         # per tensor, of max|ref_t| + 5e-3 of the largest gradient


@pytest.mark.parametrize("cfg", ["c2", "c3"])
def test_reduced_precision_mode_tolerance(cfg):
    """north_star's 'bf16 mode': ops.PRECISION = 'bf16' runs the weight gradients with bf16 operands (fp32 accumulation)
    and the node-level / LEM GEMMs with operands rounded once to tf32 (one MMA pass); the message kernel stays
    error-compensated.  Outputs and gradients against the float64 oracle at the stated looser tolerance, and the mode must
    really be active (error above the fp32-parity bar)."""
    from msmp_pde_b200 import models_gnn2D, ops, synth
    from oracle import models as om
    dev = torch.device("cuda:0")
    pde, data, meta = (synth.config_c2 if cfg == "c2" else synth.config_c3)(B=6, nx=100, seed=5)
    torch.manual_seed(1)
    model = models_gnn2D.MP_PDE_Solver2DLEMLinGated(pde, 25, 128, 6, meta["eq_variables"]).to(dev)
    torch.set_default_dtype(torch.float64)
    ref = om.MP_PDE_Solver2DLEMLinGated(pde, 25, 128, 6, meta["eq_variables"])
    ref.load_state_dict({k: v.double().cpu() for k, v in model.state_dict().items()})
    outr = ref(data)
    torch.sqrt(((outr - data.y) ** 2).sum()).backward()
    dd = data.clone().to(dev)
    prev = ops.PRECISION
    ops.PRECISION = "bf16"
    try:
        out = model(dd)
        torch.sqrt(((out - dd.y) ** 2).sum()).backward()
    finally:
        ops.PRECISION = prev
    e_out = rel_err(out, outr)
    refs = dict(ref.named_parameters())
    gscale = max(float(p.grad.abs().max()) for p in refs.values())
    worst = 0.0
    for name, p in model.named_parameters():
        r = refs[name].grad
        allow = BF16_GRAD_TOL * float(r.abs().max()) + 5e-3 * gscale
        worst = max(worst, float((p.grad.double().cpu() - r).abs().max()) / allow)
    print(f"reduced precision {cfg}: out rel err {e_out:.2e}, worst gradient error / allowance {worst:.3f}")
    assert 1e-5 < e_out < BF16_OUT_TOL
    assert worst <= 1.0


def test_prefetched_inputs_equal_direct_load():
    """GraphedTrainStep.prefetch (copy stream + staging buffers) feeds the same step as passing the graph directly."""
    from msmp_pde_b200 import models_gnn, synth
    from msmp_pde_b200.train_step import GraphedTrainStep
    dev = torch.device("cuda:0")
    pde, data, meta = synth.config_c1(B=2, nx=40, seed=1)
    _, data2, _ = synth.config_c1(B=2, nx=40, seed=2)
    losses = []
    for mode in ("direct", "prefetch"):
        torch.manual_seed(0)
        model = models_gnn.MP_PDE_Solver(pde, 25, 128, 6, {}).to(dev)
        opt = torch.optim.AdamW(model.parameters(), lr=1e-4)
        step = GraphedTrainStep(model, opt, data.clone().to(dev), warmup=1, preserve_state=True)
        pinned = [d.clone().apply(lambda t: t.pin_memory()) for d in (data, data2, data)]
        out = []
        if mode == "direct":
            for g in pinned:
                out.append(float(step(g)))
        else:
            step.prefetch(pinned[0])
            for i in range(3):
                loss = step(None)
                if i + 1 < 3:
                    step.prefetch(pinned[i + 1])
                out.append(float(loss))
        losses.append(out)
    assert losses[0] == losses[1]
    assert losses[0][0] != losses[0][1]


@pytest.mark.parametrize("n_nodes,degree", [(1 << 22, 6), (1 << 21, 16)])
def test_c5_multi_million_node_layer_slices_vs_oracle(n_nodes, degree):
    """BASELINE config 5 beyond 1 Mi nodes: one GNN_Layer(128,128,128,25,1) forward + backward on 4 Mi nodes x 6 and 2 Mi nodes
    x 16 in-neighbours (25 / 33 Mi edges: the per-edge tensors hold 3.2 / 4.3 G floats, past 32-bit element offsets).  The
    float64 oracle cannot run the whole graph in seconds, but the layer is local to a graph (messages stay inside a graph,
    InstanceNorm is per graph), so the graphs at the start, in the middle and at the END of the node / edge range are run
    through the oracle on their own and must equal the corresponding rows of the full run: outputs and input gradients."""
    from msmp_pde_b200 import layers, synth
    from oracle import models as om
    dev = torch.device("cuda:0")
    npg = 2048
    g = synth.large_graph(n_nodes, degree, topology="random", nodes_per_graph=npg, seed=5)
    torch.manual_seed(0)
    layer = layers.GNN_Layer(128, 128, 128, 25, 1).to(dev)
    ref = om.GNN_Layer(128, 128, 128, 25, 1).double()
    ref.load_state_dict({k: v.double().cpu() for k, v in layer.state_dict().items()})
    wsum = torch.randn(n_nodes, 128, generator=torch.Generator().manual_seed(1))       # d loss / d out
    t = {k: v.to(dev) for k, v in g.items()}
    x = t["x"].clone().requires_grad_(True)
    out = layer(x, t["u"], t["pos"], t["variables"], t["edge_index"], t["batch"])
    (out * wsum.to(dev)).sum().backward()
    torch.cuda.synchronize()
    n_graphs = n_nodes // npg
    src, dst = g["edge_index"]
    for gi in (0, n_graphs // 2 - 1, n_graphs - 2):
        lo, hi = gi * npg, (gi + 2) * npg                      # two neighbouring graphs
        e0, e1 = lo * degree, hi * degree                      # edges are sorted by destination, `degree` per node
        assert int(dst[e0]) == lo and int(dst[e1 - 1]) == hi - 1
        assert int(src[e0:e1].min()) >= lo and int(src[e0:e1].max()) < hi
        xs = g["x"][lo:hi].double().requires_grad_(True)
        outr = ref(xs, g["u"][lo:hi].double(), g["pos"][lo:hi].double(), g["variables"][lo:hi].double(),
                   g["edge_index"][:, e0:e1] - lo, g["batch"][lo:hi] - gi)
        (outr * wsum[lo:hi].double()).sum().backward()
        assert rel_err(out[lo:hi], outr) < OUT_TOL, gi
        assert rel_err(x.grad[lo:hi], xs.grad) < GRAD_TOL, gi
    assert bool(torch.isfinite(out).all()) and bool(torch.isfinite(x.grad).all())


@pytest.mark.parametrize("shape", [(1, 1), (37, 50), (6400, 50), (131072, 50)])
def test_sse_criterion_matches_torch_float64(shape):
    """msmp_sse_fwd / msmp_sse_bwd (the criterion of the captured step, train_helper.py:126 with float64 labels) against
    torch's float64 arithmetic: the sum to 1e-13, the float pair that rides in the gradient bucket to 2^-44, the gradient
    bit-exact for a unit seed (one rounding float64 -> float32 on both sides) and to 1e-7 for a scaled seed; bit-stable."""
    from msmp_pde_b200 import ops
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev).manual_seed(shape[0])
    pred = torch.randn(shape, device=dev, generator=gen).requires_grad_(True)
    y = torch.randn(shape, device=dev, generator=gen, dtype=torch.float64)
    tail = torch.zeros(2, device=dev)
    sse = ops.sse_loss(pred, y, tail)
    sse.backward()
    d = pred.detach().double() - y
    ref = float((d ** 2).sum())
    assert sse.dtype == torch.float64 and abs(float(sse) - ref) <= 1e-13 * ref
    assert abs(float(tail[0].double() + tail[1].double()) - float(sse)) <= 2.0 ** -44 * ref
    assert torch.equal(pred.grad, (2.0 * d).float())
    pred.grad = None
    sse2 = ops.sse_loss(pred, y, None)
    (1.7 * sse2).backward()
    assert torch.equal(sse2, sse)
    assert rel_err(pred.grad, 3.4 * d) < 1e-7


@pytest.mark.parametrize("N,tw,nvar", [(1, 25, 0), (37, 25, 1), (6400, 25, 2), (4099, 50, 0), (131072, 25, 1)])
def test_input_assembly_kernels_equal_framework_expressions(N, tw, nvar):
    """msmp_node_features / msmp_lem_inputs (one launch each) against the slice-copy expressions they replace
    (layers.NodeFeatures; I_t of models_gnn2D.py:421-433 and models_gnn.py:1357-1360): bit-identical."""
    from msmp_pde_b200 import ops
    from msmp_pde_b200.layers import pad32
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev).manual_seed(N + tw)
    u = torch.randn(N, 2 * tw, device=dev, generator=gen)
    pos_x = torch.rand(N, 1, device=dev, generator=gen)
    pos_t = torch.rand(N, 1, device=dev, generator=gen, dtype=torch.float64)
    variables = torch.rand(N, 1 + nvar, device=dev, generator=gen)
    dt64 = torch.cumsum(torch.ones(1, tw, dtype=torch.float64, device=dev) * 0.0371, dim=1)
    # node features
    ldu = pad32(2 * tw)
    upad, side = ops.node_features(u, pos_x, variables, ldu)
    upad_ref = torch.zeros(N, ldu, device=dev)
    upad_ref[:, :2 * tw] = u
    side_ref = torch.zeros(N, 8, device=dev)
    side_ref[:, 0:1] = pos_x
    side_ref[:, 1:2 + nvar] = variables
    assert torch.equal(upad, upad_ref) and torch.equal(side, side_ref)
    # two-field LEM input slab
    cols = [("static", pos_x, 0), ("time", u, 0), ("time", u, tw), ("clock",)] + [("static", variables, 1 + k) for k in range(nvar)]
    inp = ops.lem_inputs(tw, N, cols, clock=dt64.view(-1), node_t=pos_t.reshape(-1).contiguous())
    ref = torch.zeros(tw, N, 32, device=dev)
    ref[:, :, 0] = pos_x[:, 0]
    ref[:, :, 1] = u[:, :tw].t()
    ref[:, :, 2] = u[:, tw:].t()
    ref[:, :, 3] = (dt64 + pos_t).float().t()
    if nvar:
        ref[:, :, 4:4 + nvar] = variables[:, 1:]
    assert torch.equal(inp, ref)
    # one-field slab: [pos_x, u[:, t], variables]
    u1 = u[:, :tw].contiguous()
    inp1 = ops.lem_inputs(tw, N, [("static", pos_x, 0), ("time", u1, 0)] + [("static", variables, k) for k in range(1 + nvar)])
    ref1 = torch.zeros(tw, N, 32, device=dev)
    ref1[:, :, 0] = pos_x[:, 0]
    ref1[:, :, 1] = u1.t()
    ref1[:, :, 2:3 + nvar] = variables
    assert torch.equal(inp1, ref1)


def test_g2_gate_statistic_equals_framework_expression():
    """msmp_g2_fwd / msmp_g2_bwd (the G^2 gate of MP_PDE_Solver2DLEMLinG2, models_gnn2D.py:598-603) on an irregular kNN graph
    (out-degrees 0..many): forward and backward equal the float64 evaluation of the framework expression (gather, square,
    scatter-mean by source) to 1e-6, the forward also the deterministic segmented mean of the gathered squares; bit-stable."""
    from msmp_pde_b200 import synth
    from msmp_pde_b200.graph import get_topology
    from msmp_pde_b200.models_gnn2D import _G2MeanFn, _SourceMeanFn
    dev = torch.device("cuda:0")
    _, data, _ = synth.config_c3(B=5, nx=100, seed=2)
    ei, batch = data.edge_index.to(dev), data.batch.to(dev)
    N = data.x.shape[0]
    topo = get_topology(ei, batch, N)
    gen = torch.Generator(device=dev).manual_seed(3)
    t = torch.randn(N, 128, device=dev, generator=gen).requires_grad_(True)
    w = torch.randn(N, 128, device=dev, generator=gen)
    out = _G2MeanFn.apply(t, topo)
    (out * w).sum().backward()
    src, dst = topo.src.long(), topo.dst.long()
    ref = _SourceMeanFn.apply((t.detach()[src] - t.detach()[dst]) ** 2, topo)
    assert rel_err(out, ref) < 1e-6, "forward vs segmented mean of the gathered squares"      # (the kernel fuses x * x into the add)
    t64 = t.detach().double().requires_grad_(True)
    cnt = torch.bincount(src, minlength=N).clamp(min=1).double()
    ref64 = torch.zeros(N, 128, dtype=torch.float64, device=dev).index_add_(0, src, (t64[src] - t64[dst]) ** 2) / cnt[:, None]
    (ref64 * w.double()).sum().backward()
    assert rel_err(out, ref64) < 1e-6, "forward vs float64"
    assert rel_err(t.grad, t64.grad) < 1e-6, "backward vs float64"
    g1 = t.grad.clone()
    t.grad = None
    (_G2MeanFn.apply(t, topo) * w).sum().backward()
    assert torch.equal(t.grad, g1)
