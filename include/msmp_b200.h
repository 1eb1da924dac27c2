/* msmp_b200 -- C ABI of the B200-native MP-PDE / MSMP-PDE message-passing hot path.
 *
 * Plain pointers and sizes only (no torch types).  Every entry point
 *   - works on DEVICE pointers (fp32 / int32 unless stated), row-major, 16-byte aligned rows,
 *   - enqueues on the caller's stream and returns immediately (0 = MSMP_OK, negative = error;
 *     -1 bad argument, -2 CUDA launch/runtime error, -3 workspace too small),
 *   - never allocates and keeps no mutable global state (re-entrant per stream); scratch memory is
 *     passed in, sized by the matching *_workspace() query.
 *
 * What each one replaces in the reference (Leqr/MSMP-PDE; the reference has a single native FFI
 * boundary, `lem_cuda.forward/backward`, experiments/models_gnn.py:287-302 -- everything else below
 * replaces Python-level torch / PyG / torch_scatter calls on the same path):
 *
 *   msmp_linear_fwd / _wgrad     nn.Linear(+Swish) blocks: message_net_1 in per-node factorised form,
 *                                update_net_1/2, embedding / lemoutput / double MLPs, LEM gate GEMMs
 *                                (models_gnn.py:47-58,77-86,201-206,290; models_gnn2D.py:375-379)
 *   msmp_edge_fwd / _bwd         MessagePassing.propagate: gather + message() + aggr='mean'
 *                                (models_gnn.py:42,65,69-75,107,128,132-138; torch_scatter atomics)
 *   msmp_segment_reduce          torch_scatter.scatter(reduce='mean'|'sum') (models_gnn2D.py:600-601)
 *                                and the by-source gradient scatter of the backward pass
 *   msmp_instnorm_fwd / _bwd     PyG InstanceNorm (models_gnn.py:59,66,122,129) fused with the MSMP gate
 *                                blend  h = (1-tau) h + tau sw(.)  (models_gnn.py:1365-1368)
 *   msmp_lem_*                   lem_cuda.forward / lem_cuda.backward (models_gnn.py:290-292,300)
 *   msmp_decoder_*               output_mlp Conv1d -> Swish -> Conv1d + time stepping
 *                                (models_gnn.py:215-219,275-279; models_gnn2D.py:382-386,448-458)
 */
#ifndef MSMP_B200_H_
#define MSMP_B200_H_

#include <stddef.h>
#include <cuda_runtime_api.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSMP_B200_ABI_VERSION 1
int msmp_abi_version(void);

/* ---- dense node-level layers ------------------------------------------------------------------
 * Y[M, Nout] = epilogue( [A0 | A1 | A2] (K = sum ka) * Wt[K, ldw] + bias + side[M, 0:r] * Wside[r, ldw] )
 *   epilogue(z): z *= swish'(Zmul) if Zmul;  Ypre = z if Ypre;  z = swish(z) if act;  z += R if R;  Y = z.
 * ka[s] % 32 == 0, lda % 4 == 0, Nout % 4 == 0, ldw >= roundup(Nout,128), r <= 8.
 * aswish[s] != 0 applies swish to segment s while loading (a3 = swish(z3) is never materialised). */
int msmp_linear_fwd(const float* const* A, const int* lda, const int* ka, const int* aswish, int nseg,
                    const float* Wt, int ldw, const float* bias, const float* side, int lds, int r,
                    const float* Wside, const float* Zmul, int ldz, float* Ypre, int ldpre, int act,
                    const float* R, int ldr, float* Y, int ldy, int M, int Nout, cudaStream_t stream);

/* Tensor-core variant (tcgen05.mma kind::tf32, error-compensated 3xTF32, accumulators in TMEM): same contract,
 * the weight is passed as pre-split, pre-swizzled tile images Bimg[ntile][K/32][2 (hi,lo)][128 x 32 fp32]:
 * element (n, k) of a tile sits at float offset ((n>>3)*1024 + (n&7)*128 + ((((k>>2)^n)&7)<<4))/4 + (k&3);
 * hi = value rounded to tf32, lo = value - hi.  Wside has row stride ldws. */
size_t msmp_linear_tc_image_floats(int K, int Nout);
int msmp_linear_tc_fwd(const float* const* A, const int* lda, const int* ka, const int* aswish, int nseg,
                       const float* Bimg, const float* bias, const float* side, int lds, int r, const float* Wside,
                       int ldws, const float* Zmul, int ldz, float* Ypre, int ldpre, int act, const float* R,
                       int ldr, float* Y, int ldy, int M, int Nout, int mode, cudaStream_t stream);
/* mode 0: error-compensated 3xTF32 (fp32 parity); mode 1: reduced precision -- operands rounded to tf32, one MMA pass. */

/* One launch packs any number of weight blocks into tile images / plain side arrays.  jobs_dev: device array of
 * records {const float* src; float* dst; int ld, transpose; float sign; int kvalid, nvalid, nchunks, kind, ldd}
 * (msmp_pack_job_bytes() bytes each).  kind 0: dst[chunk][hi|lo][4096] images of Wt[k][n] = sign * (transpose ?
 * src[n*ld + k] : src[k*ld + n]) (zero outside kvalid x nvalid);  kind 1: dst[q*ldd + n] = sign * src[n*ld + q]. */
int msmp_pack_job_bytes(void);
int msmp_pack_run(const void* jobs_dev, int njobs, int max_chunks, cudaStream_t stream);

/* The inverse direction, once per backward pass: the weight-gradient kernels leave k-major blocks ([K][N], side and
 * bias rows) in a raw buffer; one launch writes the parameter-layout gradients (what autograd's AccumulateGrad of
 * the reference's nn.Linear / lem_cuda.backward outputs would hold, experiments/models_gnn.py:300) from records
 * {float* dst; const float* src0, *src1; int ldd, ld0, ld1, rows, cols; float sign1; int zero}
 * (msmp_unpack_job_bytes() bytes each):  dst[n*ldd + k] = src0[k*ld0 + n] (+ sign1 * src1[k*ld1 + n]),
 * n < rows, k < cols; zeros when `zero`.  max_tiles = max over jobs of ceil(rows/32) * ceil(cols/32). */
int msmp_unpack_job_bytes(void);
int msmp_unpack_run(const void* jobs_dev, int njobs, int max_tiles, cudaStream_t stream);

/* dWt[K, Nout] (+)= X[M, K]^T (swish(X) if xswish) * dY[M, Nout];
 * dWside[r (+1), Nout] (+)= [side | 1]^T * dY  (bias gradient = the implicit ones column when has_bias).
 * Deterministic: per-CTA partials over row ranges + fixed-order reduction. */
int msmp_linear_wgrad_splits(int M, int K, int Nout);
size_t msmp_linear_wgrad_workspace(int M, int K, int Nout, int nside);
int msmp_linear_wgrad(const float* X, int ldx, int K, int xswish, const float* dY, int lddy, int Nout,
                      const float* side, int lds, int r, int has_bias, float* dWt, float* dWside,
                      int accumulate, int M, void* workspace, size_t ws_bytes, cudaStream_t stream);

/* Tensor-core variant (both operands MN-major, 3xTF32, split-M partials in TMEM -> fixed-order reduce);
 * same contract as msmp_linear_wgrad, r + has_bias <= 8. */
int msmp_linear_wgrad_tc(const float* X, int ldx, int K, int xswish, const float* dY, int lddy, int Nout,
                         const float* side, int lds, int r, int has_bias, float* dWt, float* dWside, int accumulate,
                         int M, void* workspace, size_t ws_bytes, cudaStream_t stream);

/* dWt[K0 + K1, Nout] = [X | X1]^T dY with the two column blocks in different tensors (K0 % 128 == 0). */
int msmp_linear_wgrad_tc2(const float* X, int ldx, int K0, const float* X1, int ldx1, int K1, int xswish,
                          const float* dY, int lddy, int Nout, const float* side, int lds, int r, int has_bias,
                          float* dWt, float* dWside, int accumulate, int M, void* workspace, size_t ws_bytes,
                          cudaStream_t stream);

/* Second-generation weight gradient (csrc/wgrad_ws.cu): persistent warp-specialised CTAs, each owning a contiguous row
 * range for every output tile; rows arrive by bulk copies (TMA engine), side / bias gradients come out of the same MMAs.
 *   dWt[k][n] = sum_m [X0 | X1 | X2][m][k] dY[m][n],  dWside[q][n] = sum_m [side[m][side_c0 + q] (q < r) | 1][q] dY[m][n]
 * X: nseg (<= 3) column segments, kx[s] % 32 == 0 (swish applied to segment s when xswish[s]); Nout % 128 == 0;
 * side = base of a [M][lds] array with 16-byte aligned rows (lds <= 16), r + has_bias <= 8.
 * part / part_side: [S][sum kx][Nout] / [S][r + has_bias][Nout] partials, S = msmp_wgrad_ws_splits(...)
 * (msmp_wgrad_ws_workspace bytes for both, part_side directly behind part).  dWt != NULL: the partials are summed
 * in split order into dWt / dWside by a second launch; dWt == NULL: the caller sums them (k_unpack, msmp_unpack_run).
 * mode 0: error-compensated 3xTF32 (fp32 parity); mode 1: bf16 operands, fp32 accumulation.  Replaces the autograd
 * weight gradients of experiments/models_gnn.py:47-58,310-313. */
int msmp_wgrad_ws_splits(int M, int KB, int Nout, int nside);
size_t msmp_wgrad_ws_workspace(int M, int KB, int Nout, int nside);
int msmp_wgrad_ws(const float* const* X, const int* ldx, const int* kx, const int* xswish, int nseg, const float* dY,
                  int lddy, int Nout, const float* side, int lds, int side_c0, int r, int has_bias, float* part,
                  float* part_side, float* dWt, float* dWside, int accumulate, int M, int mode, cudaStream_t stream);

/* ---- edge kernels (edges sorted by destination; rowptr = CSR offsets by destination) ------------
 * forward : agg[i] = inv_deg[i] * sum_{e -> i} sw( sw(P[dst e] + Q[src e]) W2^T + b2 );  z2 (optional) keeps
 *           the second pre-activation for the backward pass.  W2t[k][n] = W2[n][k]. */
int msmp_edge_tiles(int E);
int msmp_edge_grid(int E);
size_t msmp_edge_fwd_workspace(int E);
int msmp_edge_fwd(const float* P, const float* Q, int ldpq, const int* src, const int* dst, const int* rowptr,
                  const float* inv_deg, const float* W2t, const float* b2, float* z2, float* agg, int E, int N,
                  void* workspace, size_t ws_bytes, cudaStream_t stream);
/* backward: given dagg -> dz1[E,128] (gradient at the first pre-activation), dP[i] = sum_{e->i} dz1[e],
 *           dW2[n][k], db2[n].  (dQ[j] = sum_{e from j} dz1[e] is msmp_segment_reduce over the CSC order.) */
size_t msmp_edge_bwd_workspace(int E);
int msmp_edge_bwd(const float* P, const float* Q, int ldpq, const int* src, const int* dst, const int* rowptr,
                  const float* inv_deg, const float* W2, const float* z2, const float* dagg, int lddagg,
                  float* dz1, float* dP, int lddp, float* dW2, float* db2, int E, int N, void* workspace,
                  size_t ws_bytes, cudaStream_t stream);

/* Tensor-core variants (tcgen05 3xTF32).  W2t_img / W2_img = tile images (msmp_linear_tc_fwd format, K = N = 128)
 * of the k-major W2^T (forward) and of W2 itself (backward).  The backward materialises dz2 and a1 [E,128];
 * dW2^T / db2 = msmp_linear_wgrad_tc(X = a1, dY = dz2, has_bias).  Workspace: msmp_edge_fwd_workspace(E). */
int msmp_edge_tc_fwd(const float* P, const float* Q, int ldpq, const int* src, const int* dst, const int* rowptr,
                     const float* inv_deg, const float* W2t_img, const float* b2, float* z2, float* agg, int E,
                     int N, void* workspace, size_t ws_bytes, cudaStream_t stream);
int msmp_edge_tc_bwd(const float* P, const float* Q, int ldpq, const int* src, const int* dst, const int* rowptr,
                     const float* inv_deg, const float* W2_img, const float* z2, const float* dagg, int lddagg,
                     float* dz2, float* a1, float* dz1, float* dP, int lddp, int E, int N, void* workspace,
                     size_t ws_bytes, cudaStream_t stream);

/* Warp-specialised tensor-core variants (edge_ws.cu): gather / MMA / epilogue roles overlap across 128-edge tiles and
 * the weight matrix is resident in tensor memory.  W is read as fp32 with explicit strides: the A operand of
 * D^T[m][edge] is A[m][k] = W[m * w_rs + k * w_cs]  (forward: A = W2[n][k], i.e. the message_net_2.0.weight parameter
 * itself with w_rs = 128, w_cs = 1, models_gnn.py:52-54; backward: A = W2^T, the same parameter with w_rs = 1,
 * w_cs = 128).  inv_deg_e[e] = inv_deg[dst[e]].  Same outputs and carry rules as the _tc_ entry points;
 * workspace: msmp_edge_ws_workspace(E). */
size_t msmp_edge_ws_workspace(int E);
int msmp_edge_ws_fwd(const float* P, const float* Q, int ldpq, const int* src, const int* dst, const int* rowptr,
                     const float* inv_deg, const float* W, int w_rs, int w_cs, const float* b2, float* z2, float* agg,
                     int E, int N, int no_isolated, void* workspace, size_t ws_bytes, cudaStream_t stream);
int msmp_edge_ws_bwd(const float* P, const float* Q, int ldpq, const int* src, const int* dst, const int* rowptr,
                     const float* inv_deg_e, const float* W, int w_rs, int w_cs, const float* z2, const float* dagg,
                     int lddagg, float* dz2, float* a1, float* dz1, float* dP, int lddp, int E, int N, int no_isolated,
                     void* workspace, size_t ws_bytes, cudaStream_t stream);

/* out[n, 0:128] = scale[n] * sum_{k in [ptr[n], ptr[n+1])} src[perm ? perm[k] : k, 0:128]
 * (scale == NULL -> sum; scale = 1/max(count,1) -> mean).  One warp per segment, fixed order, no atomics. */
int msmp_segment_reduce(const float* src, int lds, const int* perm, const int* ptr, const float* scale,
                        float* out, int ldo, int N, cudaStream_t stream);

/* ---- InstanceNorm (+ gate blend) ------------------------------------------------------------------
 * mode 0: out = IN(y0).   mode 1: out = (1 - s) h + s * swish(IN(y1)),  s = sigmoid(IN(y0)).
 * chunk_begin/end: graph-aligned node ranges; graph_chunk_ptr[B+1]: chunks of each graph; node_graph[N].
 * stat: [mode+1][B][2][128] (mean, rstd) -- written by fwd, read by bwd. */
size_t msmp_instnorm_workspace(int nchunks, int B);
int msmp_instnorm_fwd(const float* y0, const float* y1, int ld, const float* h, const int* chunk_begin,
                      const int* chunk_end, const int* graph_chunk_ptr, const int* node_graph, int nchunks, int B,
                      int N, int mode, float eps, float* stat, float* out, void* workspace, size_t ws_bytes,
                      cudaStream_t stream);
int msmp_instnorm_bwd(const float* dout, const float* y0, const float* y1, int ld, const float* h,
                      const float* stat, const int* chunk_begin, const int* chunk_end, const int* graph_chunk_ptr,
                      const int* node_graph, int nchunks, int B, int N, int mode, float* dy0, float* dy1, int lddy,
                      float* dh, void* workspace, size_t ws_bytes, cudaStream_t stream);
/* The same two operations in ONE launch each for batches in which every graph has 1..128 nodes (then statistics chunk
 * g is graph g; the reference's grids have 100 nodes, generate_data.py:342).  Same arithmetic and order, same bits. */
int msmp_instnorm1_fwd(const float* y0, const float* y1, int ld, const float* h, const int* chunk_begin,
                       const int* chunk_end, int B, int N, int mode, float eps, float* stat, float* out,
                       cudaStream_t stream);
int msmp_instnorm1_bwd(const float* dout, const float* y0, const float* y1, int ld, const float* h, const float* stat,
                       const int* chunk_begin, const int* chunk_end, int B, int N, int mode, float* dy0, float* dy1,
                       int lddy, float* dh, cudaStream_t stream);

/* ---- LEM recurrence (replaces lem_cuda.forward / .backward, models_gnn.py:290-292,300) -------------
 * One step t:  G[N,384] = [y_{t-1} | I_t] W^T + b         (msmp_linear_fwd)
 *              msmp_lem_gate_z: a = dt*sig(G0), b = dt*sig(G1), zc = tanh(G2), z_t = (1-b) z_{t-1} + b zc
 *              L[N,128] = [z_t | I_t] Wz^T + bz            (msmp_linear_fwd)
 *              msmp_lem_gate_y: tL = tanh(L), y_t = (1-a) y_{t-1} + a tL
 * gates[4][N][128] = (a, b, zc, tL) is the per-step state kept for the backward pass.
 * Backward step (reverse t): msmp_lem_bwd_y (dy in/out, gy external grad or NULL) -> dL, dG[:,0:128];
 * dz_tot = dz + dL Wz[:, :128] (msmp_linear_fwd); msmp_lem_bwd_z -> dG[:,128:384], dz;
 * dy += dG W[:, :128] (msmp_linear_fwd); weight gradients by msmp_linear_wgrad over all T*N rows. */
int msmp_lem_gate_z(const float* G, const float* z_prev, float dt, float* gates, float* z_new, int N,
                    cudaStream_t stream);
int msmp_lem_gate_y(const float* L, const float* y_prev, float* gates, float* y_new, int N, cudaStream_t stream);
int msmp_lem_bwd_y(float* dy, const float* gy, const float* y_prev, const float* gates, float dt, float* dL,
                   float* dG, int N, cudaStream_t stream);
int msmp_lem_bwd_z(const float* dz_tot, const float* gz, const float* z_prev, const float* gates, float dt,
                   float* dG, float* dz, int N, cudaStream_t stream);

/* ---- decoder: Conv1d(C,8,K1,stride S1) -> Swish -> Conv1d(8,C,K2) + time stepping -------------------
 * (models_gnn.py:208-224,275-279 with C = 1; models_gnn2D.py:382-391,448-458 with C = 2)
 * h[N, C*128]; out[n, c*TW + k] = base + dt[k] * diff[n, c, k], base = u[n, TW-1] (C = 1) or u[n, c*TW + k].
 * za[N, 8*L1] keeps the first pre-activation.  L1 = (128 - K1)/S1 + 1, TW = L1 - K2 + 1.
 * bwd: dh[N, C*128], dW = [w1 | b1 | w2 | b2] gradients in the parameters' own layouts (fixed-order reduce). */
int msmp_decoder_nweights(int C, int K1, int K2);
size_t msmp_decoder_bwd_workspace(int N, int C, int K1, int K2);
int msmp_decoder_fwd(const float* h, const float* w1, const float* b1, const float* w2, const float* b2,
                     const float* u, int ldu, const float* dt, float* za, float* out, int N, int C, int K1, int S1,
                     int L1, int K2, int TW, cudaStream_t stream);
int msmp_decoder_bwd(const float* dout, const float* h, const float* za, const float* w1, const float* w2,
                     const float* dt, float* dh, float* dW, int N, int C, int K1, int S1, int L1, int K2, int TW,
                     void* workspace, size_t ws_bytes, cudaStream_t stream);

/* Persistent tensor-core LEM: ALL T steps in one launch, one CTA per 64 nodes (the recurrence is per node); the
 * GEMMs are issued transposed (weights = A operand, state tile = B operand, N = 64 nodes), see csrc/lem_tc.cu.
 * inp [T][N][32] zero-padded inputs (ninp <= 8 real columns).  Npad = N rounded up to 64.  Every array is row-major.
 *   Wt_in / Wzt_in = rows 128.. of the k-major packs W^T [160 x 384] / Wz^T [160 x 128]: the input part of both
 *         affine maps, bias + inp . w_in, is evaluated in registers inside the gate epilogues
 *   Wimg / Wzimg: tile images (msmp_linear_tc_fwd format) of the STATE rows W^T[:128] / Wz^T[:128]
 *   Y, Z  [T+1][N][128] (Y[0], Z[0] = initial state; slabs 1..T are written)
 *   gates [T][Npad][512] = a | b | zc | tL  (kept for the backward)
 * Backward: Wzh_img / Wh_img = images of Wz[:, :128] ([128 x 128]) and W[:, :128] ([384 x 128]) read as k-major;
 * gY / gZ external gradients ([T][N][128], or [N][128] for t = T-1 when g_last_only; may be NULL);
 * dy / dz [Npad][128], zero on entry, gradient wrt the initial state on exit; dG [T][N][384], dL [T][N][128] feed
 * the weight-gradient GEMMs.
 * One backward launch walks the steps t = t_end-1 .. t_begin; a caller that wants to overlap the weight-gradient
 * GEMMs of finished steps with the rest of the recurrence splits [0, T) into consecutive launches, last range first
 * (dy / dz carry the state gradient between them). */
int msmp_lem_tc_fwd(const float* inp, int ninp, const float* Wt_in, const float* Wzt_in, const float* Wimg,
                    const float* Wzimg, const float* bias, const float* bias_z, float* Y, float* Z, float* gates,
                    float dt, int T, int N, int Npad, int mode, cudaStream_t stream);
int msmp_lem_tc_bwd(const float* Wzh_img, const float* Wh_img, const float* Y, const float* Z, const float* gates,
                    const float* gY, const float* gZ, int g_last_only, float* dG, float* dL, float* dy, float* dz,
                    float dt, int T, int t_begin, int t_end, int N, int Npad, int mode, cudaStream_t stream);

/* out[i] = g[i] * swish'(z[i])  (n % 4 == 0) */
int msmp_mul_dswish(const float* g, const float* z, float* out, size_t n, cudaStream_t stream);

/* ---- optimizer step of the captured training step ---------------------------------------------------------------
 * torch.optim.AdamW semantics (experiments/train.py:410; decoupled weight decay, no amsgrad / maximize) over any number
 * of tensors in ONE launch, every hyper-parameter read from DEVICE memory so that a CUDA-graph replay follows the
 * scheduler (train.py:411,437 rewrites param_group['lr'] only).  jobs_dev: records {float* p, *g, *m, *v; int n, group}
 * (msmp_adamw_job_bytes() each; m / v are the optimizer's own exp_avg / exp_avg_sq).  chunks_dev: int2 {job, first
 * element}, one per CTA, msmp_adamw_chunk() elements each.  hyper_dev: msmp_adamw_hyper_floats() floats per param group
 * = {lr, 1 - beta1, beta2, eps, weight_decay, 1 - beta1^t, sqrt(1 - beta2^t), 1 - beta2}.  gscale_dev (may be NULL): device scalar
 * every gradient is multiplied by first (and written back), e.g. 1 / (2 sqrt(SSE)) of loss = sqrt(SSE),
 * experiments/train_helper.py:126,138. */
int msmp_adamw_job_bytes(void);
int msmp_adamw_chunk(void);
int msmp_adamw_hyper_floats(void);
int msmp_adamw_run(const void* jobs_dev, const void* chunks_dev, int nchunks, const float* hyper_dev,
                   const float* gscale_dev, cudaStream_t stream);
/* loss = sqrt(sse_hi_lo[0] + sse_hi_lo[1]) (double), gscale = 0.5 / loss: the two scalars of a step whose summed squared
 * error arrives as a float pair (hi + lo) -- the last two elements of the all-reduced gradient bucket. */
int msmp_loss_scalars(const float* sse_hi_lo, double* loss, float* gscale, cudaStream_t stream);

/* Summed squared error of the training criterion and its gradient (experiments/train_helper.py:126,138:
 * MSELoss(reduction='sum') on float64 labels, loss = sqrt of it): sse = sum_i (double(pred[i]) - y[i])^2 over n elements,
 * deterministic (fixed blocks, fixed summation trees); sse_hi_lo (may be NULL) receives the same value split exactly into a
 * float pair (the two trailing elements of the gradient bucket).  workspace: msmp_sse_workspace(n) bytes.
 * msmp_sse_bwd: dpred[i] = float(2 * g * (double(pred[i]) - y[i])), g a device scalar (NULL = 1). */
size_t msmp_sse_workspace(size_t n);
int msmp_sse_fwd(const float* pred, const double* y, size_t n, void* workspace, size_t ws_bytes, double* sse,
                 float* sse_hi_lo, cudaStream_t stream);
int msmp_sse_bwd(const float* pred, const double* y, const double* g, size_t n, float* dpred, cudaStream_t stream);

/* ---- input assembly of a forward pass ------------------------------------------------------------------------------
 * msmp_node_features: the per-forward constant node inputs every layer of the stack shares (the u_i - u_j, pos_i - pos_j
 * and variables operands of message(), experiments/models_gnn.py:70-75): upad[n][0..ldu) = [u[n][0..F_u) | 0] and
 * side[n][0..8) = [pos_x[n], variables[n][0..V), 0...]; all fp32, u / variables contiguous [N,F_u] / [N,V], ldu % 4 == 0.
 * msmp_lem_inputs: the zero-padded input slab inp[t][n][0..32) of the LEM recurrence (I_t of models_gnn2D.py:421-433,
 * models_gnn.py:1357-1360); column c < ncols is described by cols[c] (HOST array): STATIC = src[n*ld + off], TIME =
 * src[n*ld + off + t] (src fp32), CLOCK = (float)(clock[t] + node_t[n]) evaluated in double (cumsum(dt)_t + pos_t). */
#define MSMP_LEM_MAX_COLS 8
#define MSMP_LEM_COL_STATIC 0
#define MSMP_LEM_COL_TIME 1
#define MSMP_LEM_COL_CLOCK 2
typedef struct msmp_lem_col {
  const void* src;
  int ld;
  int off;
  int kind;
  int reserved;
} msmp_lem_col;
int msmp_lem_inputs(const msmp_lem_col* cols, int ncols, const double* clock, const double* node_t, int T, int N,
                    float* inp, cudaStream_t stream);
int msmp_node_features(const float* u, int F_u, const float* pos_x, const float* variables, int V, int N, float* upad,
                       int ldu, float* side, cudaStream_t stream);

/* ---- G^2 gate statistic (MP_PDE_Solver2DLEMLinG2, experiments/models_gnn2D.py:598-603) ------------------------------
 * out[s] = inv[s] * sum over the out-edges e of node s, in CSC order, of (t[s] - t[dst[e]])^2 -- what
 * torch_scatter.scatter((x_i - x_j)^2, edge_index[0], reduce='mean') computes, without the [E,128] intermediate and
 * without atomics; rows of 128 floats; src / dst in CSR (destination-sorted) edge order, csc_perm = CSR edge ids sorted by
 * source, colptr / rowptr the matching offsets, inv[s] = 1 / max(out-degree, 1).  msmp_g2_bwd: dt = d out / d t applied to g. */
int msmp_g2_fwd(const float* t, const int* colptr, const int* csc_perm, const int* dst, const float* inv, float* out, int N,
                cudaStream_t stream);
int msmp_g2_bwd(const float* t, const float* g, const int* colptr, const int* csc_perm, const int* src, const int* dst,
                const int* rowptr, const float* inv, float* dt, int N, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MSMP_B200_H_ */
