/*
 * trajsde_b200.h — C ABI of the B200-native fused Euler–Maruyama solve that replaces TrajSDE's two solver call sites.
 *
 * Reference interfaces each entry point replaces (paths relative to the reference repo daeheepark/TrajSDE):
 *   trajsde_euler_fwd   : torchsde.sdeint(sde, y0, ts, dt=.., method='euler')      models/decoders/dec_hivt_nusargo_sde.py:88
 *                         sdeint_dual(sde, y0, ts, nus_mask, dt=..)                models/utils/sdeint.py:110-197,
 *                           called at models/encoders/enc_hivt_nusargo_sde_sep2.py:149 and :274
 *                         (solver loop models/utils/sdeint.py:326-384, Euler step :467-485, f/g glue :537-566,
 *                          nets dec…sde.py:107-127,141-158,180-195 / enc…sep2.py:372-398,412-440,462-482)
 *   trajsde_euler_bwd   : torch.autograd through that solver (config `adjoint: false`, yml:41): discretise-then-optimise
 *   trajsde_enc_fwd     : the encoder recurrence 21 x [sdeint_dual one step + GRU_Unit jump]
 *                           enc_hivt_nusargo_sde_sep2.py:128-182 + models/utils/ode_utils.py:136-152
 *   trajsde_enc_bwd     : torch.autograd through that recurrence
 *   trajsde_gru_fwd/bwd : GRU_Unit.forward as its own operator           models/utils/ode_utils.py:136-152 (enc…sep2.py:165-169)
 *   trajsde_philox_dw   : BrownianInterval increments W(t1)-W(t0) ~ N(0,(t1-t0) I)  models/utils/sdeint.py:983-984
 *
 * Conventions
 *   - Every pointer named *device* is a CUDA device pointer owned by the caller; the library never allocates or frees
 *     device memory and keeps no reference after the call returns.  Scratch comes from `workspace` (size queried with
 *     trajsde_*_workspace_bytes).  All calls only ENQUEUE work on `cuda_stream` (a cudaStream_t / CUstream passed as void*;
 *     NULL = legacy default stream) and never synchronise the device.
 *   - All tensors are fp32, dim (channels) == 64, rows are independent (one row = one agent latent, or agent x mode).
 *   - Return value: 0 on success, negative TrajsdeStatus otherwise; trajsde_last_error_string() gives a thread-local
 *     message.  The library never calls abort()/exit() and never throws across the ABI.
 *   - Re-entrant; safe from several host threads on different streams.
 */
#ifndef TRAJSDE_B200_H_
#define TRAJSDE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TRAJSDE_ABI_VERSION 9
#define TRAJSDE_DIM 64

typedef enum {
  TRAJSDE_OK = 0,
  TRAJSDE_ERR_INVALID_ARGUMENT = -1, /* null pointer, bad size, struct_bytes mismatch */
  TRAJSDE_ERR_UNSUPPORTED = -2,      /* dim != 64, unknown mode, misaligned pointer */
  TRAJSDE_ERR_WORKSPACE = -3,        /* workspace too small */
  TRAJSDE_ERR_CUDA = -4,             /* CUDA runtime / launch error */
  TRAJSDE_ERR_NO_DEVICE = -5         /* no sm_100 device / kernel image unavailable */
} TrajsdeStatus;

/* Arithmetic mode of the MLP contractions (state, h, dW, accumulation and the Euler update are fp32 in every mode). */
typedef enum {
  TRAJSDE_MODE_EXACT_F32 = 0, /* fp32 FFMA on CUDA cores, libm-accurate tanh/sigmoid: the validation path            */
  TRAJSDE_MODE_TC_F16 = 1,    /* tcgen05 tensor cores, fp16 operands (10-bit mantissa = TF32 precision), fp32 accum,
                                 MUFU tanh; the throughput path                                                     */
  TRAJSDE_MODE_TC_BF16 = 2    /* tcgen05 tensor cores, bf16 operands (reserved)                                      */
} TrajsdeMode;

/* One 3-layer MLP exactly as stored by nn.Linear (row-major [out,in]); device pointers.
 * drift:     w1[64,66] b1[64] w2[64,64] b2[64] w3[64,64] b3[64]     (FFunc.net[0],[2],[4])
 * diffusion: w1[64,66] b1[64] w2[64,64] b2[64] w3[1,64]  b3[1]      (GFunc.net[0],[2],[4]; sigmoid applied on top)
 * Columns 64 and 65 of w1 multiply sin(t0) and cos(t0) (time features, dec…sde.py:124-126). */
typedef struct {
  const float* w1;
  const float* b1;
  const float* w2;
  const float* b2;
  const float* w3;
  const float* b3;
} TrajsdeMlp;

/* Same layout for gradients (device, written — not accumulated — by the backward call). */
typedef struct {
  float* w1;
  float* b1;
  float* w2;
  float* b2;
  float* w3;
  float* b3;
} TrajsdeMlpGrad;

/* Step schedule, computed on the host by replaying the reference's float32 time loop (models/utils/sdeint.py:340-384)
 * and uploaded once per (ts, dt); device pointers.
 *   step_tab[4*k + {0,1,2,3}] = t0_k, h_k = t1_k - t0_k, sin(t0_k), cos(t0_k)          k = 0..n_steps-1
 *   out_begin[k] .. out_begin[k+1]-1 = the outputs j (0-based, ys index j+1) whose interpolation interval ends with
 *                  step k; n_steps+1 entries, non-decreasing, out_begin[n_steps] == n_outputs
 *   out_w[2*j + {0,1}] = w0_j, w1_j  with  ys[j+1] = w0_j * Y[k] + w1_j * Y[k+1]  (torchsde linear_interp)            */
typedef struct {
  int32_t n_steps;
  int32_t n_outputs;
  const float* step_tab;
  const int32_t* out_begin;
  const float* out_w;
} TrajsdeSchedule;

/* Brownian increments: either caller-supplied (validation / parity) or generated in-kernel.
 *   dw != NULL : dw[n_steps, rows, 64] contiguous, slab k consumed by schedule step k (Var = h_k).
 *   dw == NULL : Philox4x32-10, key = (seed lo, seed hi), counter = (global_row lo, global_row hi, step_offset + k,
 *                channel/4); Box–Muller; scaled by sqrt(h_k).  global_row = row + row_offset so results do not depend
 *                on how rows are sharded over GPUs. */
typedef struct {
  const float* dw;
  uint64_t seed;
  uint64_t row_offset;
  uint32_t step_offset;
  uint32_t reserved;
  const uint64_t* seed_dev; /* device pointer or NULL: the effective key is seed + *seed_dev — a seed that lives in device memory lets
                               a captured CUDA graph draw fresh noise on every replay (bump the word between replays) */
} TrajsdeNoise;

typedef struct {
  uint32_t struct_bytes; /* = sizeof(TrajsdeEulerFwdArgs): ABI check */
  int32_t mode;          /* TrajsdeMode */
  int64_t rows;
  int32_t dim;           /* must be 64 */
  int32_t flags;         /* reserved, 0 */
  TrajsdeSchedule sched;
  TrajsdeMlp drift;
  TrajsdeMlp diffusion;     /* rows with alt_mask[row] != 0, or all rows when alt_mask == NULL  (decoder g / encoder g_nus) */
  TrajsdeMlp diffusion_alt; /* rows with alt_mask[row] == 0 (encoder g_argo); ignored when alt_mask == NULL              */
  const uint8_t* alt_mask;  /* device [rows] (torch.bool storage) or NULL                                               */
  TrajsdeNoise noise;
  const float* y0;       /* [rows,64], row stride y0_row_stride elements, unit channel stride */
  int64_t y0_row_stride;
  float* ys;             /* n_outputs+1 slabs: ys[0] = y0; element (t,row,c) at t*ys_t_stride + row*ys_row_stride + c */
  int64_t ys_t_stride;
  int64_t ys_row_stride;
  float* g_last;         /* [rows] diffusion evaluated at the start of the LAST step (sdeint_dual's 2nd result), or NULL */
  float* states;         /* [n_steps, rows, 64] state at the START of every step (saved for backward), or NULL           */
  void* workspace;
  int64_t workspace_bytes;
} TrajsdeEulerFwdArgs;

/* bits of *status (TrajsdeEulerBwdArgs / TrajsdeEncBwdArgs): the scaled adjoint left the range in which its fp16 delta operands
 * are exact to 2^-11 (it grew by more than ~2^17 over max|incoming gradient|): gradients of this call may be clipped — rerun with
 * TRAJSDE_BWD_FLAG_EXACT_KERNELS or TRAJSDE_MODE_EXACT_F32. */
#define TRAJSDE_STATUS_ADJOINT_RANGE 1
/* trajsde_enc_bwd only: a CTA of the single-launch sweep waited ~30 s for a tile another CTA of the same launch had to hand over (the
 * launch was not fully co-resident, e.g. the device is shared through MPS); the call's results are invalid — rerun with
 * TRAJSDE_BWD_FLAG_PER_STEP_LAUNCHES. */
#define TRAJSDE_STATUS_SWEEP_TIMEOUT 2

/* TrajsdeEulerBwdArgs.flags */
#define TRAJSDE_BWD_FLAG_EXACT_KERNELS 1 /* run the fp32 CUDA-core backward even in TC_F16 mode (A/B validation) */
/* Skip the rows whose incoming gradients (grad_ys, grad_g_last) are all zero: their adjoint is zero at every step, so they contribute
 * exactly nothing to grad_y0 (written as 0) or to any weight gradient.  One extra scan of grad_ys finds them on the device (no host
 * synchronisation); pays off under a winner-takes-all loss such as the reference's L2 (losses/L2.py:17-20: one of the 10 modes of an actor
 * receives a gradient).  Honoured by the tensor-core kernels with a single diffusion net; ignored otherwise. */
#define TRAJSDE_BWD_FLAG_SKIP_ZERO_ROWS 2
/* trajsde_enc_bwd: run the reverse sweep as 2 launches per iteration of the recurrence (GRU backward, SDE-step backward) instead of ONE
 * persistent launch whose CTAs take the GRU / SDE roles and hand tiles to each other through global progress counters (A/B validation;
 * also the path for devices where the launch cannot be fully co-resident). */
#define TRAJSDE_BWD_FLAG_PER_STEP_LAUNCHES 4

/* Backward kernels by mode: EXACT_F32 -> fp32 CUDA-core dgrad sweep + wgrad (euler_bwd_exact.cu).  TC_F16 with a single
 * diffusion net -> fused tensor-core dgrad+wgrad (euler_bwd_tc.cu; fp16 operands, fp32 accumulation, adjoint carried with a
 * power-of-two loss scale chosen from max|grad|).  TC_F16 with alt_mask (dual diffusion) -> the fp32 kernels. */
typedef struct {
  uint32_t struct_bytes;
  int32_t mode;
  int64_t rows;
  int32_t dim;
  int32_t flags;           /* TRAJSDE_BWD_FLAG_* */
  TrajsdeSchedule sched;
  TrajsdeMlp drift;
  TrajsdeMlp diffusion;
  TrajsdeMlp diffusion_alt;
  const uint8_t* alt_mask;
  TrajsdeNoise noise;
  const float* states;       /* [n_steps, rows, 64] from the forward call */
  const float* grad_ys;      /* dL/d ys, same indexing as ys (incl. slab 0), or NULL */
  int64_t grad_ys_t_stride;
  int64_t grad_ys_row_stride;
  const float* grad_g_last;  /* [rows] dL/d g_last or NULL */
  float* grad_y0;            /* [rows,64] contiguous */
  TrajsdeMlpGrad grad_drift;
  TrajsdeMlpGrad grad_diffusion;
  TrajsdeMlpGrad grad_diffusion_alt; /* required iff alt_mask != NULL */
  int32_t* status;           /* device int32 or NULL: the TC kernels OR in TRAJSDE_STATUS_* bits (never cleared by the library) */
  const float* grad_amax;    /* device float or NULL; only read together with row_flags: an estimate of max |grad_ys| over the flagged rows
                                (within a few binades: it picks a power-of-two loss scale that has 2^17 of head-room) — the call then reads
                                no gradient before the sweep itself.  trajsde_heads_bwd writes exactly this (TrajsdeHeadsBwdArgs.grad_amax) */
  const uint8_t* row_flags;  /* device [rows] or NULL.  With TRAJSDE_BWD_FLAG_SKIP_ZERO_ROWS: row_flags[r] == 0 PROMISES that every incoming
                                gradient of row r is zero (its grad_ys entries are then never read and may be uninitialised), so the call
                                skips its own activity scan — trajsde_heads_bwd produces exactly these flags */
  void* workspace;
  int64_t workspace_bytes;
} TrajsdeEulerBwdArgs;

int trajsde_abi_version(void);
const char* trajsde_last_error_string(void);

/* Number of resident CTAs the persistent kernels use on the current device (148 SMs on B200 x CTAs/SM); <0 on error. */
int trajsde_device_sm_count(void);

int64_t trajsde_euler_fwd_workspace_bytes(int32_t mode, int64_t rows, int32_t n_steps, int32_t dual_diffusion);
int trajsde_euler_fwd(const TrajsdeEulerFwdArgs* args, void* cuda_stream);

/* Upper bound over both backward kernel families of `mode` (valid whatever `flags` the call then uses). */
int64_t trajsde_euler_bwd_workspace_bytes(int32_t mode, int64_t rows, int32_t n_steps, int32_t dual_diffusion);
int trajsde_euler_bwd(const TrajsdeEulerBwdArgs* args, void* cuda_stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Fused encoder recurrence: n_steps x [one Euler step of the dual-diffusion SDE + GRU_Unit jump], the loop body of
 * LocalEncoderSDESepPara2.forward (models/encoders/enc_hivt_nusargo_sde_sep2.py:128-182) around sdeint_dual
 * (models/utils/sdeint.py:110-197) and GRU_Unit.forward (models/utils/ode_utils.py:136-152).  The latent state of a
 * 128-row tile stays on chip for all iterations.  Tensor-core mode only (TRAJSDE_MODE_TC_F16).
 * ------------------------------------------------------------------------------------------------------------------ */

/* GRU_Unit parameters, nn.Linear layout (ode_utils.py:115-133): update_gate / reset_gate / new_state_net = Linear(128,64),
 * Tanh, Linear(64,64)(, Sigmoid).  Columns 0..63 of u1/r1 multiply h_cur, 64..127 the input; columns 0..63 of n1 multiply
 * the input, 64..127 reset*h_cur (ode_utils.py:137,142). */
typedef struct {
  const float* u1; const float* ub1; const float* u2; const float* ub2;   /* update_gate   [64,128] [64] [64,64] [64] */
  const float* r1; const float* rb1; const float* r2; const float* rb2;   /* reset_gate                                */
  const float* n1; const float* nb1; const float* n2; const float* nb2;   /* new_state_net                             */
} TrajsdeGru;

typedef struct {
  uint32_t struct_bytes;
  int32_t mode;              /* TRAJSDE_MODE_TC_F16 */
  int64_t rows;
  int32_t dim;               /* 64 */
  int32_t flags;
  TrajsdeSchedule sched;     /* one step per loop iteration: step_tab[4*i] = t0_i, h_i, sin t0_i, cos t0_i (n_steps <= 32);
                                out_* tables unused */
  TrajsdeMlp drift;
  TrajsdeMlp diffusion;      /* g_nus  (rows with alt_mask != 0, or all rows if alt_mask == NULL) */
  TrajsdeMlp diffusion_alt;  /* g_argo (rows with alt_mask == 0) */
  const uint8_t* alt_mask;   /* device [rows] nus_mask or NULL */
  TrajsdeGru gru;
  TrajsdeNoise noise;        /* dw[n_steps, rows, 64] or Philox (counter step = step_offset + iteration) */
  const float* h0;           /* [rows,64] initial hidden state, row stride h0_row_stride elements */
  int64_t h0_row_stride;
  const float* aa_out;       /* [n_slots, rows, 64] GRU inputs (output of the AA encoder, enc…sep2.py:107-121) */
  const int32_t* slot;       /* device [n_steps]: data slot consumed by iteration i (run_backwards: 20, 19, ..., 0) */
  const uint8_t* obs_mask;   /* device [rows, n_slots] bool: actors_mask (enc…sep2.py:100); row stride obs_mask_row_stride */
  int64_t obs_mask_row_stride;
  float* latent;             /* out [n_steps, rows, 64]: post-GRU state of every iteration (latent_ys, :180,184) */
  float* g_out;              /* out [n_steps, rows]: diffusion evaluated at the start of every iteration's step (:149,171) */
  float* y1_out;             /* out [n_steps, rows, 64] or NULL: state after the SDE step, before the GRU jump (saved for
                                trajsde_enc_bwd) */
  void* workspace;
  int64_t workspace_bytes;
} TrajsdeEncFwdArgs;

int64_t trajsde_enc_fwd_workspace_bytes(int32_t mode, int64_t rows, int32_t n_steps, int32_t dual_diffusion);
int trajsde_enc_fwd(const TrajsdeEncFwdArgs* args, void* cuda_stream);

/* Gradients of the GRU_Unit parameters (device, written — not accumulated). */
typedef struct {
  float* u1; float* ub1; float* u2; float* ub2;
  float* r1; float* rb1; float* r2; float* rb2;
  float* n1; float* nb1; float* n2; float* nb2;
} TrajsdeGruGrad;

/* Backward of the whole encoder recurrence (what autograd computes through enc…sep2.py:128-182 with `adjoint: false`): one call,
 * a reverse sweep over the iterations that chains, per iteration, the GRU_Unit backward (fp32) and the one-step dual-diffusion
 * SDE backward (fused tensor-core dgrad+wgrad, one pass per diffusion net).  Consumes what trajsde_enc_fwd produced with
 * y1_out != NULL: latent (post-GRU states) and y1 (pre-GRU states); activations are recomputed. */
typedef struct {
  uint32_t struct_bytes;
  int32_t mode;              /* TRAJSDE_MODE_TC_F16 */
  int64_t rows;
  int32_t dim;               /* 64 */
  int32_t flags;
  TrajsdeSchedule sched;     /* as in TrajsdeEncFwdArgs (step_tab only) */
  TrajsdeMlp drift;
  TrajsdeMlp diffusion;
  TrajsdeMlp diffusion_alt;
  const uint8_t* alt_mask;   /* device [rows] nus_mask or NULL (single diffusion net) */
  TrajsdeGru gru;
  TrajsdeNoise noise;        /* the forward call's noise: dw[n_steps, rows, 64] or the same Philox seed / offsets */
  const float* h0;           /* [rows,64] contiguous */
  const float* aa_out;       /* [n_slots, rows, 64] */
  int32_t n_slots;
  int32_t reserved;
  const int32_t* slot;       /* device [n_steps] */
  const uint8_t* obs_mask;
  int64_t obs_mask_row_stride;
  const float* latent;       /* [n_steps, rows, 64] from the forward call */
  const float* y1;           /* [n_steps, rows, 64] from the forward call (y1_out) */
  const float* grad_latent;  /* dL/d latent [n_steps, rows, 64] or NULL */
  const float* grad_g;       /* dL/d g_out [n_steps, rows] or NULL */
  float* grad_h0;            /* out [rows,64] */
  float* grad_aa_out;        /* out [n_slots, rows, 64]: slabs of the visited slots are written, the caller zero-fills the rest; or NULL */
  TrajsdeMlpGrad grad_drift;
  TrajsdeMlpGrad grad_diffusion;
  TrajsdeMlpGrad grad_diffusion_alt; /* required iff alt_mask != NULL */
  TrajsdeGruGrad grad_gru;
  int32_t* status;           /* as in TrajsdeEulerBwdArgs */
  void* workspace;
  int64_t workspace_bytes;
} TrajsdeEncBwdArgs;

int64_t trajsde_enc_bwd_workspace_bytes(int32_t mode, int64_t rows, int32_t n_steps, int32_t dual_diffusion);
int trajsde_enc_bwd(const TrajsdeEncBwdArgs* args, void* cuda_stream);

/* Stand-alone GRU_Unit jump (models/utils/ode_utils.py:136-152) for hosts that keep the reference's encoder loop: forward
 * h_next = mask ? (1-u) n + u h_cur : h_cur, and its backward (what autograd computes through GRU_Unit.forward).  Tensor cores,
 * fp16 operands / fp32 accumulation; hidden = input = units = 64 only. */
typedef struct {
  uint32_t struct_bytes;
  int32_t mode;              /* TRAJSDE_MODE_TC_F16 */
  int64_t rows;
  int32_t dim;               /* 64 */
  int32_t flags;
  TrajsdeGru gru;
  const float* h_cur;        /* [rows,64] contiguous */
  const float* x;            /* [rows,64] contiguous: input_tensor */
  const uint8_t* mask;       /* [rows] bool */
  float* h_next;             /* forward out [rows,64] */
  const float* grad_h_next;  /* backward in  [rows,64] */
  float* grad_h_cur;         /* backward out [rows,64] */
  float* grad_x;             /* backward out [rows,64] */
  TrajsdeGruGrad grad_gru;   /* backward out */
  int32_t* status;           /* reserved (NULL) */
  void* workspace;
  int64_t workspace_bytes;
} TrajsdeGruArgs;

int64_t trajsde_gru_workspace_bytes(int32_t mode, int64_t rows);
int trajsde_gru_fwd(const TrajsdeGruArgs* args, void* cuda_stream);
int trajsde_gru_bwd(const TrajsdeGruArgs* args, void* cuda_stream);

/* Fused decoder heads over the solver outputs: for every (row, t) latent x[t][row][0..63]
 *     out[h][row][t][0..1] = W2_h . relu(LayerNorm(W1_h x + b1_h)) + b2_h          h = 0 (self.decoder), 1 (self.scale)
 * replaces the two nn.Sequential calls of SDEDecoder.forward (models/decoders/dec_hivt_nusargo_sde.py:50-61, 96, 98); the ELU,
 * +1 and +min_scale of :98-99 stay with the caller (they act on the 2-channel result).  Forward only (inference); fp16 operands /
 * fp32 accumulation for the 64x64 layers, fp32 LayerNorm and projections; dim = 64 only. */
typedef struct {
  const float* w1;   /* [64,64] net[0].weight */
  const float* b1;   /* [64]    net[0].bias */
  const float* ln_g; /* [64]    net[1].weight */
  const float* ln_b; /* [64]    net[1].bias */
  const float* w2;   /* [2,64]  net[3].weight */
  const float* b2;   /* [2]     net[3].bias */
} TrajsdeHead;

/* TrajsdeHeadsArgs.flags / TrajsdeHeadsBwdArgs.flags: write (read the gradient of) the decoder's RESULT tensor instead of the two raw
 * head outputs — out['loc'] = cat(loc, elu(scale) + 1.0 + min_scale) of dec_hivt_nusargo_sde.py:98-100, [rows, n_t, 4] contiguous at
 * out[0] (grad_out[0]); needs n_heads == 2.  Deletes the ELU, the two adds and the cat (and their backward) from the caller's graph. */
#define TRAJSDE_HEADS_FLAG_CAT4 1

typedef struct {
  uint32_t struct_bytes;
  int32_t mode;              /* TRAJSDE_MODE_TC_F16 */
  int64_t rows;
  int32_t dim;               /* 64 */
  int32_t flags;             /* TRAJSDE_HEADS_FLAG_* */
  int32_t n_t;               /* time slabs of x (60 for the reference decoder) */
  int32_t n_heads;           /* 1 (uncertain = False) or 2 */
  TrajsdeHead head[2];
  float ln_eps;              /* nn.LayerNorm eps (1e-5) */
  float min_scale;           /* TRAJSDE_HEADS_FLAG_CAT4: self.min_scale of the decoder (yml: 1e-3) */
  const float* x;            /* element (t, row, c) at x + t * x_t_stride + row * x_row_stride + c; 16-byte aligned, strides % 4 == 0 */
  int64_t x_row_stride;
  int64_t x_t_stride;
  float* out[2];             /* per head [rows, n_t, 2] contiguous; CAT4: out[0] = [rows, n_t, 4], out[1] unused */
  void* workspace;
  int64_t workspace_bytes;
} TrajsdeHeadsArgs;

int64_t trajsde_heads_workspace_bytes(int32_t mode);
int trajsde_heads_fwd(const TrajsdeHeadsArgs* args, void* cuda_stream);

/* Backward of the fused decoder heads (training): what autograd computes through the two nn.Sequential heads of SDEDecoder.forward
 * (dec_hivt_nusargo_sde.py:50-61, 96, 98) — dL/dx summed over the heads and the gradients of every head parameter, from
 * dL/dout_h [rows, n_t, 2].  Only the (point, head) pairs with a non-zero dL/dout are processed (the reference's winner-takes-all L2
 * loss, losses/L2.py:12-20, leaves ~5 % of them): grad_x must be ZERO-FILLED by the caller (or see row_flags), rows of active points are
 * accumulated into.
 * fp32 arithmetic on CUDA cores in every mode (the validation-grade twin of the tensor-core forward). */
typedef struct {
  float* w1; float* b1; float* ln_g; float* ln_b; float* w2; float* b2;   /* written, not accumulated */
} TrajsdeHeadGrad;

typedef struct {
  uint32_t struct_bytes;
  int32_t mode;              /* any TrajsdeMode (the arithmetic is fp32) */
  int64_t rows;
  int32_t dim;               /* 64 */
  int32_t flags;             /* TRAJSDE_HEADS_FLAG_* */
  int32_t n_t;
  int32_t n_heads;           /* 1 or 2 */
  TrajsdeHead head[2];
  float ln_eps;
  float min_scale;           /* unused (the ELU derivative needs only the recomputed raw scale) */
  const float* x;            /* as in TrajsdeHeadsArgs */
  int64_t x_row_stride;
  int64_t x_t_stride;
  const float* grad_out[2];  /* per head dL/dout [rows, n_t, 2] contiguous, or NULL (no gradient reaches that head); CAT4: grad_out[0] =
                                dL/d out['loc'] [rows, n_t, 4] (channels 2..3 go through the ELU derivative), grad_out[1] unused */
  float* grad_x;             /* element (t, row, c) at grad_x + t * gx_t_stride + row * gx_row_stride + c; zero-filled by the caller */
  int64_t gx_row_stride;
  int64_t gx_t_stride;
  TrajsdeHeadGrad grad_head[2];
  float* grad_amax;          /* device float or NULL (row_flags mode): receives max |value written to grad_x| (0 when nothing was active) */
  uint8_t* row_flags;        /* device [rows] or NULL.  Non-NULL: grad_x may be UNINITIALISED on entry; the call writes row_flags[r] = "some
                                point of row r carries a gradient", zero-fills the n_t entries of exactly those rows and accumulates into
                                them — rows with flag 0 are left untouched (hand the flags to trajsde_euler_bwd, which then never reads them) */
  void* workspace;
  int64_t workspace_bytes;
} TrajsdeHeadsBwdArgs;

int64_t trajsde_heads_bwd_workspace_bytes(int32_t mode);
int trajsde_heads_bwd(const TrajsdeHeadsBwdArgs* args, void* cuda_stream);

/* ------------------------------------------------------------------------------------------------------------------
 * The decoder stage's prologue and the two training losses around the solve; fp32, fused, one call each.
 * ------------------------------------------------------------------------------------------------------------------ */

/* aggr_embed: hidden_0[m * N + n] = ReLU(LayerNorm(W [global_embed[m, n] ; local_embed[n]] + b))
 * replaces `self.aggr_embed(torch.cat((global_embed, local_embed.expand(num_modes, ...)), dim=-1))` + the view to [modes * N, 64]
 * (models/decoders/dec_hivt_nusargo_sde.py:26-29, 82-85).  Forward writes `out`; backward (what autograd computes through it) reads
 * grad_out and writes grad_global, grad_local (summed over the modes) and the parameter gradients (written, not accumulated). */
typedef struct {
  uint32_t struct_bytes;
  int32_t n_modes;
  int64_t n_actors;
  const float* global_embed; /* [n_modes, n_actors, 64] contiguous */
  const float* local_embed;  /* [n_actors, 64] contiguous */
  const float* w;            /* [64, 128] aggr_embed[0].weight: columns 0..63 multiply global_embed, 64..127 local_embed */
  const float* b;            /* [64] */
  const float* ln_g;         /* [64] aggr_embed[1].weight */
  const float* ln_b;         /* [64] aggr_embed[1].bias */
  float ln_eps;
  float reserved;
  float* out;                /* forward: [n_modes * n_actors, 64] */
  const float* grad_out;     /* backward: dL/dout */
  float* grad_global;        /* backward out [n_modes, n_actors, 64] */
  float* grad_local;         /* backward out [n_actors, 64] */
  float* grad_w; float* grad_b; float* grad_ln_g; float* grad_ln_b;
  void* workspace;           /* backward only */
  int64_t workspace_bytes;
} TrajsdeAggrArgs;

/* pi head: pi[n, m] = w2 . ReLU(LayerNorm(W1 [local_embed[n] ; global_embed[m, n]] + b1)) + b2
 * replaces `self.pi(torch.cat((local_embed.expand(num_modes, ...), global_embed), dim=-1)).squeeze(-1).t()`
 * (models/decoders/dec_hivt_nusargo_sde.py:63-67, 92-94).  Forward only: no loss of the reference configuration reads pi. */
typedef struct {
  uint32_t struct_bytes;
  int32_t n_modes;
  int64_t n_actors;
  const float* global_embed; /* [n_modes, n_actors, 64] contiguous */
  const float* local_embed;  /* [n_actors, 64] contiguous */
  const float* w1;           /* [64, 128] pi[0].weight: columns 0..63 multiply local_embed, 64..127 global_embed */
  const float* b1;           /* [64] */
  const float* ln_g;         /* [64] pi[1].weight */
  const float* ln_b;         /* [64] pi[1].bias */
  const float* w2;           /* [64] pi[3].weight */
  const float* b2;           /* [1]  pi[3].bias */
  float ln_eps;
  float reserved;
  float* out;                /* [n_actors, n_modes] */
} TrajsdePiArgs;

int trajsde_pi_head_fwd(const TrajsdePiArgs* args, void* cuda_stream);

int64_t trajsde_aggr_embed_workspace_bytes(int64_t n_modes, int64_t n_actors);
int trajsde_aggr_embed_fwd(const TrajsdeAggrArgs* args, void* cuda_stream);
int trajsde_aggr_embed_bwd(const TrajsdeAggrArgs* args, void* cuda_stream);

/* L2 (losses/L2.py:10-27): per actor the mode with the smallest masked mean displacement wins; loss = mean over the valid
 * (actor, slot) pairs of that mode's ||y - loc||; 0 when nothing is valid.  Forward writes loss, count (= reg_mask.sum()) and
 * best_mode; backward writes dL/dloc for the winning mode's valid slots into the ZERO-FILLED grad_loc. */
typedef struct {
  uint32_t struct_bytes;
  int32_t n_modes;
  int64_t n_actors;
  int32_t n_t;
  int32_t reserved;
  const float* loc;          /* element (m, n, t, c) at loc + ((m * n_actors + n) * n_t + t) * loc_stride + c, c = 0, 1 */
  int64_t loc_stride;        /* 4 for the reference's cat(loc, scale) output, 2 for a plain loc tensor */
  const float* target;       /* [n_actors, n_t, 2] data['y'] */
  const uint8_t* reg_mask;   /* [n_actors, n_t] bool */
  float* loss;               /* [1] */
  float* count;              /* [1] number of valid (actor, slot) pairs, as float */
  int32_t* best_mode;        /* [n_actors] */
  const float* grad_loss;    /* backward: [1] dL/dloss */
  float* grad_loc;           /* backward: same indexing as loc with grad_loc_stride; zero-filled by the caller */
  int64_t grad_loc_stride;
  void* workspace;
  int64_t workspace_bytes;
} TrajsdeL2Args;

int64_t trajsde_l2_loss_workspace_bytes(int64_t n_actors);
int trajsde_l2_loss_fwd(const TrajsdeL2Args* args, void* cuda_stream);
int trajsde_l2_loss_bwd(const TrajsdeL2Args* args, void* cuda_stream);

/* DiffBCE (losses/diff_BCE.py:11-16 with the labels of enc…sep2.py:194-195): loss = BCE(diff_in, 0) + BCE(diff_out, 1), mean
 * reduction, log terms clamped at -100 like nn.BCELoss; grad_in / grad_out (optional) receive dloss/d diff_*. */
typedef struct {
  uint32_t struct_bytes;
  int32_t reserved;
  int64_t n_in;
  int64_t n_out;
  const float* diff_in;
  const float* diff_out;
  float* loss;               /* [1] */
  float* grad_in;            /* [n_in] or NULL */
  float* grad_out;           /* [n_out] or NULL */
  void* workspace;
  int64_t workspace_bytes;
} TrajsdeBceArgs;

int64_t trajsde_diff_bce_workspace_bytes(void);
int trajsde_diff_bce(const TrajsdeBceArgs* args, void* cuda_stream);

/* Materialise the in-kernel Brownian increments: dw_out[n_steps, rows, 64] = exactly what trajsde_euler_fwd would draw
 * with the same TrajsdeNoise (dw field ignored) and schedule.  Lets parity tests replay Philox runs through the oracle. */
int trajsde_philox_dw(const TrajsdeSchedule* sched, const TrajsdeNoise* noise, int64_t rows, float* dw_out,
                      void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* TRAJSDE_B200_H_ */
