"""msmp_wgrad_ws (csrc/wgrad_ws.cu) against float64 math: every operand shape the models use, ragged row counts, the
split-M partial mode, and the bf16 operand mode with its stated tolerance."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _always_the_persistent_kernel():
    """ops.wgrad_use_ws() hands small / very tall products to k_wgrad_tc; these tests are about k_wgrad_ws itself."""
    from msmp_pde_b200 import ops
    prev = (ops.WGRAD_WS_MIN_ROWS, ops.WGRAD_WS_MAX_TALL_ROWS)
    ops.WGRAD_WS_MIN_ROWS, ops.WGRAD_WS_MAX_TALL_ROWS = 0, 1 << 62
    yield
    ops.WGRAD_WS_MIN_ROWS, ops.WGRAD_WS_MAX_TALL_ROWS = prev

from tests.util import rel_err  # noqa: E402

# (M, K0, K1, Nout, r, has_bias, xswish)  -- K1 = 0: single segment
SHAPES = [
    (6400, 128, 0, 128, 0, True, True),          # update_net_2 (X = swish(z3))
    (6400, 128, 128, 128, 3, True, False),       # update_net_1: [h | agg], side = variables
    (37632, 128, 0, 128, 0, True, False),        # message_net_2 over edges
    (6400, 128, 64, 256, 4, True, False),        # P | Q projection: [h | upad], side = [pos, variables]
    (6400, 128, 32, 256, 2, True, False),        # P | Q projection, one field (tw 25 -> 32 columns)
    (6400, 128, 128, 256, 4, True, False),       # tw 50, two fields: F_u = 100 -> 128 columns (one dY block per CTA)
    (25 * 300, 128, 32, 384, 0, True, False),    # LEM G map: [y | inputs], three dY blocks
    (25 * 300, 128, 32, 128, 0, True, False),    # LEM L map
    (1000, 32, 0, 128, 0, True, False),          # embedding Linear (27 -> 32 zero padded inputs)
    (1000, 128, 0, 256, 0, True, False),         # double_mlp
    (17, 128, 0, 128, 0, True, False),           # fewer rows than one chunk + ragged tail
    (131, 128, 128, 128, 3, False, False),       # no bias row
    (4099, 128, 64, 256, 0, False, False),       # no side block at all
]


def _ref(X, X1, dY, side, r, has_bias, xswish):
    Xd = X.double()
    if xswish:
        Xd = Xd * torch.sigmoid(Xd)
    if X1 is not None:
        Xd = torch.cat([Xd, X1.double()], 1)
    dW = Xd.t() @ dY.double()
    cols = []
    if side is not None and r:
        cols.append(side[:, :r].double())
    if has_bias:
        cols.append(torch.ones(dY.shape[0], 1, dtype=torch.float64, device=dY.device))
    dWs = torch.cat(cols, 1).t() @ dY.double() if cols else None
    return dW, dWs


def _inputs(M, K0, K1, Nout, r, seed):
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(seed)
    X = torch.randn(M, K0, device=dev, generator=g)
    X1 = torch.randn(M, K1, device=dev, generator=g) if K1 else None
    dY = torch.randn(M, Nout, device=dev, generator=g)
    side8 = torch.randn(M, 8, device=dev, generator=g)
    side = side8[:, 1:] if r == 3 else side8           # a column-sliced view, like layers.py passes for update_net_1
    return X, X1, dY, (side if r else None)


@pytest.mark.parametrize("shape", SHAPES)
def test_wgrad_ws_fp32_parity(shape):
    from msmp_pde_b200 import ops
    M, K0, K1, Nout, r, has_bias, xswish = shape
    X, X1, dY, side = _inputs(M, K0, K1, Nout, r, seed=M + Nout)
    assert ops.WGRAD_WS and ops.PRECISION == "fp32"
    dW, dWs = ops.linear_wgrad(X, dY, xswish=xswish, side=side, r=r, has_bias=has_bias, X1=X1)
    rW, rWs = _ref(X, X1, dY, side, r, has_bias, xswish)
    assert rel_err(dW, rW) < 5e-6
    if rWs is not None:
        assert rel_err(dWs, rWs) < 5e-6
    # run-to-run bit identity (fixed-order reduction, no atomics)
    dW2, dWs2 = ops.linear_wgrad(X, dY, xswish=xswish, side=side, r=r, has_bias=has_bias, X1=X1)
    assert torch.equal(dW, dW2) and (dWs is None or torch.equal(dWs, dWs2))


def test_wgrad_ws_partials_sum_to_the_gradient():
    from msmp_pde_b200 import ops
    from msmp_pde_b200._lib import lib
    M, K0, K1, Nout, r = 6400, 128, 64, 256, 4
    X, X1, dY, side = _inputs(M, K0, K1, Nout, r, seed=1)
    S = lib.msmp_wgrad_ws_splits(M, K0 + K1, Nout, r + 1)
    assert S > 1
    part = torch.full((S, K0 + K1, Nout), float("nan"), device=X.device)
    part_s = torch.full((S, r + 1, Nout), float("nan"), device=X.device)
    ops.linear_wgrad(X, dY, side=side, r=r, has_bias=True, X1=X1, dWt=part, dWside=part_s)
    dW, dWs = ops.linear_wgrad(X, dY, side=side, r=r, has_bias=True, X1=X1)
    acc, acc_s = torch.zeros_like(dW), torch.zeros_like(dWs)
    for s in range(S):                    # the order k_reduce_partials2 / k_unpack use
        acc += part[s]
        acc_s += part_s[s]
    assert torch.equal(acc, dW) and torch.equal(acc_s, dWs)


BF16_TOL = 1e-2          # of max|ref|: bf16 operands (8-bit mantissa), fp32 accumulation over >= 1000 rows


@pytest.mark.parametrize("shape", SHAPES[:8])
def test_wgrad_ws_bf16_mode(shape):
    from msmp_pde_b200 import ops
    M, K0, K1, Nout, r, has_bias, xswish = shape
    X, X1, dY, side = _inputs(M, K0, K1, Nout, r, seed=7)
    prev = ops.PRECISION
    ops.PRECISION = "bf16"
    try:
        dW, dWs = ops.linear_wgrad(X, dY, xswish=xswish, side=side, r=r, has_bias=has_bias, X1=X1)
    finally:
        ops.PRECISION = prev
    rW, rWs = _ref(X, X1, dY, side, r, has_bias, xswish)
    assert rel_err(dW, rW) < BF16_TOL
    assert rel_err(dW, rW) > 1e-5            # really ran with bf16 operands
    if rWs is not None:
        assert rel_err(dWs, rWs) < BF16_TOL


def test_kernel_selection_policy():
    """Small products go to k_wgrad_tc, the rest to msmp_wgrad_ws (k_wgrad_ws; k_wgrad_ts for at most 192 operand columns,
    which also takes the very tall LEM products; a tall product that k_wgrad_ts cannot take stays on k_wgrad_tc); the kernels
    give the same gradient to fp32 accuracy."""
    from msmp_pde_b200 import ops
    ops.WGRAD_WS_MIN_ROWS, ops.WGRAD_WS_MAX_TALL_ROWS = 1 << 16, 1 << 18
    assert not ops.wgrad_use_ws(6400, 256, 128, 4)
    assert ops.wgrad_use_ws(131072, 256, 128, 4)
    assert ops.wgrad_use_ws(25 * 131072, 160, 384, 1) == ops.WGRAD_TS          # LEM dG: k_wgrad_ts
    assert ops.wgrad_use_ws(520192, 128, 128, 1) == ops.WGRAD_TS               # edge dW2: k_wgrad_ts
    assert not ops.wgrad_use_ws(25 * 131072, 192, 128, 1)                      # 192 + side block > 192 columns: k_wgrad_tc
    assert ops.wgrad_use_ws(25 * 131072, 256, 128, 1)
    # k_wgrad_ts (tall rule) against k_wgrad_tc and float64 on a strided three-block dY (tensor-map copies), ragged row count
    ops.WGRAD_WS_MAX_TALL_ROWS = 1 << 16
    X, X1, dY, side = _inputs(70001, 128, 32, 384, 0, seed=5)
    assert ops.wgrad_use_ws(70001, 160, 384, 1) == ops.WGRAD_TS
    a, a_s = ops.linear_wgrad(X, dY, has_bias=True, X1=X1)
    ops.WGRAD_WS_MIN_ROWS = 1 << 30
    b, b_s = ops.linear_wgrad(X, dY, has_bias=True, X1=X1)
    ops.WGRAD_WS_MIN_ROWS, ops.WGRAD_WS_MAX_TALL_ROWS = 1 << 16, 1 << 18
    rW, rWs = _ref(X, X1, dY, None, 0, True, False)
    assert rel_err(a, rW) < 5e-6 and rel_err(a_s, rWs) < 5e-6
    # k_wgrad_tc keeps a row range's whole sum in the tensor-core accumulator, whose truncating adds make the error grow like
    # rows^1.5 per CTA (1e-5 of max|ref| here; scripts/wgrad_error_growth.py) -- the reason the tall products are not its
    assert rel_err(b, rW) < 3e-5 and rel_err(b_s, rWs) < 5e-6
    X, X1, dY, side = _inputs(70000, 128, 64, 256, 4, seed=3)
    a, a_s = ops.linear_wgrad(X, dY, side=side, r=4, has_bias=True, X1=X1)              # k_wgrad_ws
    ops.WGRAD_WS_MIN_ROWS = 1 << 30
    b, b_s = ops.linear_wgrad(X, dY, side=side, r=4, has_bias=True, X1=X1)              # k_wgrad_tc
    assert rel_err(a, b) < 5e-6 and rel_err(a_s, b_s) < 5e-6


@pytest.mark.parametrize("M", [70001, 520192, 25 * 131072])
def test_wgrad_ts_error_does_not_grow_with_rows(M):
    """k_wgrad_ts reads the tensor-core accumulator out every 8 chunks of 32 rows into a round-to-nearest fp32 sum, so its error
    stays at the 2e-6 level of a short product however tall the operand is (LEM weight gradients of the C4 batch: 3.3 Mi rows,
    where a single tensor-core accumulation chain per CTA is off by 4.6e-4 of max|ref| and an FFMA GEMM by 5e-6)."""
    from msmp_pde_b200 import ops
    if not ops.WGRAD_TS:
        pytest.skip("MSMP_WGRAD_TS=0")
    Nout = 128 if M == 520192 else 384
    X, X1, dY, side = _inputs(M, 128, 32, Nout, 0, seed=5)
    a, a_s = ops.linear_wgrad(X, dY, has_bias=True, X1=X1)
    rW, rWs = _ref(X, X1, dY, None, 0, True, False)
    assert rel_err(a, rW) < 5e-6
    assert rel_err(a_s, rWs) < 1e-5          # plain fp32 column sums of up to 67 Ki rows per CTA
