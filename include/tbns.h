/*
 * tbns.h — C ABI of the B200-native Physics-Attention path (libtbns.so, sm_100a only).
 *
 * The reference (OnurBasci/TransformerBasedNavierStokeSolver) is pure PyTorch and has no FFI of its
 * own; each entry point below replaces a span of reference Python and cites it (paths relative to
 * the reference root).  Conventions:
 *   - every pointer is a DEVICE pointer unless named h_*; the caller allocates all outputs,
 *     saved-for-backward buffers and workspaces (sizes via the *_bytes helpers); the library never
 *     frees or retains caller memory beyond the call;
 *   - row-major, fp32 unless stated; `stream` is a cudaStream_t passed as void*; calls are
 *     asynchronous w.r.t. the host and re-entrant;
 *   - return 0 on success, negative error code otherwise; tbns_last_error() gives the message
 *     (thread-local).  There is NO CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef TBNS_H
#define TBNS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TBNS_OK 0
#define TBNS_ERR_INVALID (-1)
#define TBNS_ERR_CUDA (-2)
#define TBNS_ERR_UNSUPPORTED (-3)

/* operand precision of the large token-dimension contractions (projections, to_out, MLP) */
#define TBNS_PREC_FP32 0 /* fp32 operands, fp32 accumulate       (north_star "fp32 mode", <=1e-5): 3xTF32 split products
                            on tcgen05 (error <= 2^-21 per product) where the shape fills a tensor-core tile, fp32 FMA otherwise */
#define TBNS_PREC_BF16 1 /* bf16 operands, fp32 accumulate       (north_star "bf16 mode", <=2e-3) */
#define TBNS_PREC_FP32_EXACT 2 /* fp32 operands, fp32 FMA accumulate on the SIMT engine for every shape (also: env TBNS_FP32_TC=0) */

const char* tbns_last_error(void);
int tbns_version(void);
/* 1 when the runtime sees an sm_100 device; compute entry points fail otherwise. */
int tbns_device_ok(void);
/* multiProcessorCount of the current device (grid-sizing heuristics of the host side) */
int tbns_sm_count(void);

/* ---------------------------------------------------------------------------------------------
 * Generic strided GEMM with the gathers/epilogues the path needs:  C[b] = epi(A[b] (M x K) * B[b] (K x N))
 * Building block of every dense contraction on the path (conv-as-implicit-GEMM, Linear, deslice+to_out,
 * all dgrad/wgrad).  See DESIGN.md §kernels.
 * ------------------------------------------------------------------------------------------- */
typedef struct tbns_gemm_desc {
  int M, N, K;
  int batch;                    /* >=1 ; batch strides in elements                                   */
  long long sA, sB, sC, sR, sAux;
  const float* A; long long lda;
  int a_kind;                   /* 0: A(m,k)=A[m*lda+k] (K contiguous) ; 1: A(m,k)=A[k*lda+m]        */
  const float* B; long long ldb;
  int b_kind;                   /* 0: B(k,n)=B[n*ldb+k] (K contiguous) ; 1: B(k,n)=B[k*ldb+n]        */
  float* C; long long ldc;
  /* 3x3/pad-1 gather on the token index of A (row-major grid, token = i*Wg + j, per batch image):
   * 0 none; 1: a_kind 0, m = token, k = tap*Cin + ci (conv fprop / dgrad with flip=1);
   * 2: a_kind 1, k = token, m = tap*Cin + ci (conv wgrad)                                          */
  int conv_mode, Hg, Wg, Cin, flip;
  const float* bias;            /* [N] or NULL                                                       */
  const float* residual; long long ldr; /* [M,N] or NULL                                            */
  int act;                      /* 0 none ; 1 GELU(erf), pre-activation -> aux_out if non-NULL ;
                                   2 multiply by GELU'(aux_in[m,n])                                  */
  float* aux_out; const float* aux_in; long long ldaux;
  int precision;                /* TBNS_PREC_*                                                       */
  int split_k; float* ws;       /* split_k>1: ws holds split_k*batch*M*N floats                      */
  /* scatter==1: C is ignored; element (m=tap*Cin+ci, n) goes to (n<I ? Cx : Cfx)[(n%I)*Cin*taps + ci*taps + tap]
   * i.e. straight into nn.Conv2d / nn.Linear weight.grad layout                                    */
  int scatter, I, taps; float* Cx; float* Cfx;
} tbns_gemm_desc;

int tbns_gemm(const tbns_gemm_desc* d, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Tensor-core path of the projections (tcgen05.mma + TMA + TMEM, bf16 operands, fp32 accumulate):
 *   C[m, n] = sum_{tap,ci} A[shift(m,tap), ci] * W[n, tap*Cin + ci] (+ bias[n]),  m = token of a [Bimg,Hg,Wg] grid
 *   A_bf16 [Bimg*Hg*Wg, Cin] bf16 (NHWC == the reference's [B,N,C]); W_bf16 [N, taps*Cin] bf16; C fp32 [.., ldc].
 *   taps=9: 3x3/pad-1 conv pair in_project_x|in_project_fx (model/Physics_Attention.py:94-97) as one GEMM;
 *   flip=1: its transposed convolution (dgrad); taps=1: nn.Linear (:36-39; use Hg=1, Wg=#tokens).
 * Shapes: Cin % 64 == 0, N % 64 == 0 (tbns_gemm_tc_supported); other shapes go through tbns_gemm.
 * ------------------------------------------------------------------------------------------- */
typedef struct tbns_tc_desc {
  const void* A16;              /* bf16 activations [Bimg*Hg*Wg, Cin]                                  */
  int Bimg, Hg, Wg, Cin, taps, flip;
  const void* W16;              /* bf16 weights [w_batched ? Bimg : 1][N][taps*Cin]                     */
  int N, w_batched;             /* w_batched=1: image b uses its own weight matrix (deslice: P[b])      */
  const float* bias;            /* [N] or NULL                                                          */
  int act;                      /* 0 / 1 / 2 as tbns_gemm_desc.act ; 3: GELU, aux_out receives GELU'(pre) instead of pre ;
                                   4: multiply by aux_in[m,n] itself (the derivative stored by act 3): the backward
                                   contraction then needs no transcendental                               */
  float* aux_out; const float* aux_in; long long ldaux;
  const float* residual; long long ldr;
  float* C; long long ldc;      /* fp32 output or NULL                                                  */
  void* C16; long long ldc16;   /* bf16 output or NULL (operand of the next tensor-core contraction)    */
  int round_tf32;               /* 1: round the fp32 output to TF32 (RNA): it feeds kind::tf32 MMAs downstream */
  int aux_bf16;                 /* 1: aux_out / aux_in point to bf16 buffers (halves the GELU side-stream traffic)   */
  /* Fused LayerNorm of the OUTPUT rows (ln_gamma != NULL; N in {128, 256} so that one tile holds whole rows; needs C with
   * ldc == N): ln_out16[m, :] = bf16(LayerNorm(C[m, :]) * ln_gamma + ln_beta), ln_mean / ln_rstd [rows] for the backward.
   * This is the NEXT stage's nn.LayerNorm (model/Transolver_Structured_Mesh_2D.py:70-71: ln_1 / ln_2 read the residual
   * stream the contraction has just produced) - it costs no extra pass over HBM.                                          */
  const float* ln_gamma; const float* ln_beta; void* ln_out16; float* ln_mean; float* ln_rstd; float ln_eps;
} tbns_tc_desc;
int tbns_gemm_tc_supported(int Cin, int N, int taps);
int tbns_gemm_tc(const tbns_tc_desc* d, void* stream);

/* Token-contraction ("wgrad") tensor-core GEMM, both operands token-major (MN-major UMMA descriptors):
 *   D[(tap, a), n] = sum_token A[shift(token, tap), a] * B[token, n]
 * conv/linear weight gradients (SURVEY §8 a-bwd: dW_x = corr(x, dX), dW_fx = corr(x, dF)), MLP weight gradients and
 * dP = w^T dOut (batched=1: one output per image).  fp32 partials of `split_k` token ranges go to ws
 * (split_k * batch * taps*Ma * Nb floats) and are reduced in a fixed order; the result is written to C
 * ([batch][taps*Ma][Nb], ldc/sC) or, with scatter=1, straight into the Conv2d/Linear weight.grad layout
 * (see tbns_gemm_desc.scatter).  Shapes: Ma % 128 == 0, Nb % 64 == 0. */
typedef struct tbns_tc_wgrad_desc {
  const void* A16; int Ma;      /* bf16 [Bimg*Hg*Wg, Ma]                                                */
  const void* B16; int Nb;      /* bf16 [Bimg*Hg*Wg, Nb]                                                */
  int Bimg, Hg, Wg, taps, batched;
  int split_k; float* ws;
  float* C; long long ldc, sC;
  int scatter, I; float* Cx; float* Cfx;
} tbns_tc_wgrad_desc;
int tbns_gemm_tc_wgrad_supported(int Ma, int Nb, int taps);
int tbns_gemm_tc_wgrad(const tbns_tc_wgrad_desc* d, void* stream);
/* fp32 -> bf16 (round-to-nearest-even) copy; in/out 16-byte aligned */
int tbns_cast_bf16(const float* in, void* out, long long n, void* stream);

/* ---------------------------------------------------------------------------------------------
 * LayerNorm over the last dim (nn.LayerNorm(hidden_dim), eps 1e-5):
 *   model/Transolver_Structured_Mesh_2D.py:59,63,66 and forward :70-73
 * ------------------------------------------------------------------------------------------- */
/* y (fp32) and/or y16 (bf16 copy: TMA operand of the next tensor-core contraction); either may be NULL */
int tbns_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* y, void* y16, float* mean, float* rstd,
                       int rows, int C, float eps, void* stream);
/* dx = LN'(dy) + (dres ? dres : 0); dgamma/dbeta reduced deterministically through ws
 * (tbns_layernorm_bwd_ws_floats(C) floats). */
size_t tbns_layernorm_bwd_ws_floats(int C);
int tbns_layernorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma,
                       const float* dres, float* dx, void* dx16 /* optional bf16 copy of dx */, float* dgamma, float* dbeta,
                       float* dsum /* optional [C]: column sums of dx = bias gradient of the layer that produced this stream */,
                       float* ws, int rows, int C, void* stream);

/* same for a bf16 incoming gradient dy16 (written bf16-only by the data-gradient GEMM in bf16 mode); C in {128, 256, 512};
 * sums [3][C] = dgamma | dbeta | column sums of dx */
int tbns_layernorm_bwd_supported16(int C);
/* sums may be NULL: the per-CTA partials stay in ws as [tbns_layernorm_bwd_ctas(rows)][3*C] for the caller to reduce
 * (tbns_reduce_rows), e.g. on another stream */
int tbns_layernorm_bwd_ctas(int rows);
int tbns_layernorm_bwd16(const void* dy16, const float* x, const float* mean, const float* rstd, const float* gamma,
                         const float* dres, float* dx, void* dx16, float* sums, float* ws, int rows, int C, void* stream);

/* Last layer with out_dim = 1: out = mlp2(ln_3(x)) = LN(x) . w + b in one pass over x
 * (model/Transolver_Structured_Mesh_2D.py:66-67,72-73).  C in {128, 256, 512}. */
int tbns_ln_linear1_supported(int C);
int tbns_ln_linear1_fwd(const float* x, const float* gamma, const float* beta, const float* w /* [C] */, const float* b /* [1] */,
                        float* out /* [rows] */, float* mean, float* rstd, int rows, int C, float eps, void* stream);
/* same with the scalar of row r written to out[r*ldo]: the prediction lands straight in column t of a frame history
 * [rows, ldo], so a closed-loop rollout needs no concatenation (model/SOL_Transolver_Structured_Mesh_2D.py:47-52,
 * ns_vorticity_unrolling.py:269-277) */
int tbns_ln_linear1_fwd_strided(const float* x, const float* gamma, const float* beta, const float* w, const float* b, float* out,
                                long long ldo, float* mean, float* rstd, int rows, int C, float eps, void* stream);
/* dx = LN'(dout (x) w) + dres.  sums [3][C]: S_c = sum_r dout[r]*xhat[r][c] | D = sum_r dout[r] (replicated over c) |
 * column sums of dx; then dgamma = w*S, dbeta = w*D, dw = gamma*S + beta*D, db = D.  ws: tbns_layernorm_bwd_ws_floats(C). */
int tbns_ln_linear1_bwd(const float* dout /* [rows] */, const float* w, const float* x, const float* mean, const float* rstd,
                        const float* gamma, const float* dres, float* dx, void* dx16, float* sums, float* ws, int rows, int C,
                        void* stream);

/* ---------------------------------------------------------------------------------------------
 * Weight packing for the projections (nn.Conv2d 3x3 / nn.Linear pair in_project_x, in_project_fx:
 * model/Physics_Attention.py:18-19, :74-75).  taps = 9 (structured) or 1 (irregular).
 *   Wf [2I][taps*C]  : Wf[n][tap*C+ci]     = W_{x|fx}[n%I][ci][tap]    (fprop B operand, K contiguous)
 *   Wd [C][taps*2I]  : Wd[ci][tap*2I + n]  = same element               (dgrad B operand, K contiguous)
 *   bcat [2I]
 * ------------------------------------------------------------------------------------------- */
int tbns_pack_proj_weights(const float* Wx, const float* bx, const float* Wfx, const float* bfx, float* Wf, float* Wd,
                           float* bcat, int I, int C, int taps, void* stream);
/* same with optional bf16 copies Wf16 / Wd16 (the tensor-core operands) written by the same launch; Wf / Wd may be NULL
 * when only the bf16 copies are wanted (bf16 mode: one launch per layer refreshes the operands after an optimizer step) */
int tbns_pack_proj_weights16(const float* Wx, const float* bx, const float* Wfx, const float* bfx, float* Wf, float* Wd, void* Wf16,
                             void* Wd16, float* bcat, int I, int C, int taps, void* stream);
/* nn.Linear weight W [R][K] fp32 -> bf16 copy out16 [R][Kp] and / or transposed copy outT16 [Kp][R] (columns K..Kp-1 zero):
 * the K-major operands of the forward (y = x W^T) and data-gradient (dx = dy W) contractions, one launch */
int tbns_cast_bf16_pair(const float* W, void* out16, void* outT16, int R, int K, int Kp, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Input packing of the `preprocess` MLP (model/Transolver_Structured_Mesh_2D.py:203-207: pos.repeat + cat(x, fx)) fused
 * with the rollout window shift (SOL_...py:47-52): writes the bf16 operand [rows, Kp] of the first Linear in one pass.
 *   out16[r][0:R] = tab16[r % N][0:R] (bf16 table, e.g. the unified-position features) | src1[r*ld1 + j], j < F1 |
 *   src2[r*ld2 + j], j < F2 (fp32 sources, arbitrary row strides: a window of a frame history) | zeros up to Kp.
 * ------------------------------------------------------------------------------------------- */
int tbns_pack_inputs(const void* tab16, int R, const float* src1, long long ld1, int F1, const float* src2, long long ld2, int F2,
                     void* out16, int Kp, long long rows, int N, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Slice stage (model/Physics_Attention.py:40-42 / :98-101):
 *   w[b,n,h,g] = softmax_g((X[b,n,h,:].Ws[g,:] + bs[g]) / tau_h),  partial sums over token chunks of
 *   s = sum_n w and Tt = sum_n w (x) F.   XF = [X | F] is [B*N, 2I].  clamp=1 clamps tau to [0.1,5].
 *   part: [B,H,groups,G,D+1] with groups = tbns_slice_groups(B,N,H); w fp32 and/or w16 bf16 (either may be NULL).
 *   dim_head D in {8,16,32,64}, slice_num G in {4,8,16,32,64}.
 * ------------------------------------------------------------------------------------------- */
int tbns_slice_groups(int B, int N, int H);  /* partials per (batch, head) written by the slice kernels */
int tbns_pa_slice_fwd(const float* XF, const float* Ws, const float* bs, const float* temperature, float* w, void* w16,
                      float* part, int B, int N, int H, int D, int G, int clamp, void* stream);

/* Tensor-core variant of the slice stage (tcgen05.mma kind::f16 with bf16 operands on 128-token x one-head tiles, softmax
 * fused between the two contractions); bf16 mode, dim_head == 32 and slice_num in {32, 64} (tbns_pa_slice_tc_supported).
 * XF16: the projections [B*N, 2*H*D] in bf16 (the projection GEMM's C16 output - no fp32 copy of XF exists on this route).
 * Same outputs as tbns_pa_slice_fwd with w16 only. */
int tbns_pa_slice_tc_supported(int D, int G);
int tbns_pa_slice_fwd_tc(const void* XF16, const float* Ws, const float* bs, const float* temperature, void* w16, float* part,
                         int B, int N, int H, int D, int G, int clamp, void* stream);
/* backward twin: dXF16 (bf16) only; the projection-bias gradients follow from dbs and s (tbns_pa_proj_bias_grad), so no
 * dbcat_part is produced.  dw16: the deslice gradient [B,N,H*G] in bf16 (written by the dw GEMM's bf16 output) */
int tbns_pa_slice_bwd_tc(const void* XF16, const float* Ws, const float* bs, const float* temperature, const void* dw16,
                         const float* dTt, const float* ds, void* dXF16, float* dWs_part, float* dtau_part, int B, int N,
                         int H, int D, int G, int clamp, void* stream);
/* projection-bias gradients of the tensor-core route from token-reduced quantities (fixed summation order):
 *   db_x[h*D+d] = sum_g (sum_{b,chunk} dWs_part[b,h,chunk,g,D]) * Ws[g,d],   db_fx[h*D+d] = sum_{b,g} s[b,h,g] * dTt[b,h,g,d]
 * (model/Physics_Attention.py:94-97: the biases of in_project_x / in_project_fx).  dWs_part is tbns_pa_slice_bwd_tc's output. */
int tbns_pa_proj_bias_grad(const float* dWs_part, const float* Ws, const float* s, const float* dTt, float* dbx, float* dbfx,
                           int B, int H, int D, int G, int groups, void* stream);

/* Token stage (model/Physics_Attention.py:43-52 / :102-111) + fold of to_out into P:
 *   reduces `part` -> s[B,H,G], Tt[B,H,G,D]; tok = Tt/(s+1e-5); q,k,v; A = softmax(q k^T D^-1/2); O = A v;
 *   P[b,h*G+g,c] = sum_d O[b,h,g,d] * Wo[c,h*D+d].   Wo is [Cout, H*D].
 *   Optional bf16 copies for the tensor-core deslice: P16 [B,H*G,Cout] and PT16 [B,Cout,H*G] (NULL to skip). */
int tbns_pa_token_attn_fwd(const float* part, int nchunk, const float* Wq, const float* Wk, const float* Wv,
                           const float* Wo, float* s, float* Tt, float* tok, float* q, float* k, float* v, float* A,
                           float* O, float* P, void* P16, void* PT16, int B, int H, int D, int G, int Cout, void* stream);

/* Backward of the token stage.  dP [B,H*G,Cout] ->  dTt [B,H,G,D], ds [B,H,G], and per-(b,h) partials
 *   dWqkv_part [B*H,3,D,D],  dWo_part [B,Cout,H*D]  (reduce over the leading dim with tbns_reduce_rows). */
int tbns_pa_token_attn_bwd(const float* dP, const float* Wq, const float* Wk, const float* Wv, const float* Wo,
                           const float* s, const float* tok, const float* q, const float* k, const float* v,
                           const float* A, const float* O, float* dTt, float* ds, float* dWqkv_part, float* dWo_part,
                           int B, int H, int D, int G, int Cout, void* stream);

/* Backward of the slice stage (SURVEY §8 a-bwd): recomputes logits / softmax from XF.
 *   dw [B,N,H*G] (deslice gradient), dTt, ds  ->  dXF [B*N,2I] (fp32 and/or bf16 copy dXF16; either may be NULL),
 *   dWs_part [B*H*groups, G, D+1] (last column = dbs), dtau_part [B*H*groups],
 *   dbcat_part [B*groups, H, 2, D] (bias gradients of in_project_x | in_project_fx), groups = tbns_slice_groups. */
int tbns_pa_slice_bwd(const float* XF, const float* Ws, const float* bs, const float* temperature, const float* dw,
                      const float* dTt, const float* ds, float* dXF, void* dXF16, float* dWs_part, float* dtau_part,
                      float* dbcat_part, int B, int N, int H, int D, int G, int clamp, void* stream);
/* dtau_part [B,H,nchunk] -> dtemperature [H], applying the clamp mask [0.1<=tau<=5] when clamp=1 */
int tbns_pa_dtau_finish(const float* dtau_part, const float* temperature, float* dtemperature, int B, int H,
                        int nchunk, int clamp, void* stream);

/* out[j] = sum_i in[i*cols + j], i < rows (fixed order, deterministic). */
int tbns_reduce_rows(const float* in, float* out, int rows, long long cols, void* stream);

/* AdamW step over flat fp32 buffers (parameters, gradients, first / second moments, n elements each; torch.optim.AdamW
 * semantics, exp_ns.py:172-173,208-209).  hp: 7 floats on the device written by the caller before the launch
 * {lr, beta1, beta2, eps, weight_decay, 1 - beta1^t, 1 - beta2^t} - a graph replay picks up the scheduler's new values. */
int tbns_adamw_flat(float* p, const float* g, float* m, float* v, const float* hp, long long n, void* stream);
/* column sums of a [rows, cols] matrix (bias gradients); ws: tbns_colsum_ws_floats(cols) floats */
size_t tbns_colsum_ws_floats(long long cols);
int tbns_colsum_bf16(const void* in16, long long ld, float* out, float* ws, int rows, int cols, void* stream);
int tbns_colsum(const float* in, long long ld, float* out, float* ws, int rows, int cols, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TBNS_H */
