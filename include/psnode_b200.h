/*
 * psnode_b200.h -- C ABI of the B200-native fixed-grid neural ODE/DAE integrator.
 *
 * The reference (xxh0523/Py_PSNODE @ d366e75) has NO native / FFI boundary: its hot path is a Python
 * loop (neural_dae/my_solvers.py:52-131) calling Python nn.Modules.  This header is therefore the
 * boundary a maintainer would bind *underneath* the reference's Python call surface; every entry
 * point cites the reference code whose work it takes over.  The binding itself (ctypes, because the
 * reference is Python) is py_psnode_b200/_native.py and is shown in INTEGRATION.md.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch / C++ types.
 *   - every pointer is a DEVICE pointer unless the function name ends in _host.
 *   - all floating point data is IEEE fp32 (the reference runs nn.Linear default dtype everywhere).
 *   - functions return PSNODE_OK (0) or a negative PSNODE_E* code; they never throw, never allocate
 *     device memory (the caller owns outputs and workspace), are stream-ordered on `stream`
 *     (a cudaStream_t passed as void*) and never synchronise the device.
 *   - time series are passed as strided views: element (j, b, c) of a series lives at
 *     p[j*st + b*sb + c] (unit stride over the feature index c).  This honours the reference's
 *     `x.permute(1,0,2)` views of batch-major storage (neural_00_ODE_01_no_encode.py:82-84)
 *     without a copy.
 */
#ifndef PSNODE_B200_H
#define PSNODE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSNODE_ABI_VERSION 5
#define PSNODE_MAX_LAYERS 8

/* status codes */
#define PSNODE_OK 0
#define PSNODE_EINVAL (-1)      /* inconsistent dimensions / null pointers                      */
#define PSNODE_EUNSUPPORTED (-2) /* legal problem this build has no kernel for                   */
#define PSNODE_EWORKSPACE (-3)  /* workspace pointer null or too small                           */
#define PSNODE_ECUDA (-4)       /* a CUDA runtime call failed (see psnode_last_cuda_error)       */
#define PSNODE_ENODEVICE (-5)   /* no sm_100 device                                              */

/* integration scheme: neural_dae/my_fixed_grid.py:12-59 (Euler :15-18, Midpoint :23-32, RK4 3/8-rule :38-59) */
#define PSNODE_EULER 0
#define PSNODE_MIDPOINT 1
#define PSNODE_RK4 2

/* problem kind: integrate_ODE (my_solvers.py:52-80) or integrate_DAE (my_solvers.py:82-131) */
#define PSNODE_ODE 0
#define PSNODE_DAE 1

/* kernel selection for psnode_forward / psnode_backward (`impl` field) */
#define PSNODE_IMPL_AUTO 0     /* fastest kernel that supports the problem: TC8 / WIDE, else FUSED, else GENERIC */
#define PSNODE_IMPL_GENERIC 1  /* shared-memory-weight kernel, reference formulation of layer 1  */
#define PSNODE_IMPL_FUSED 2    /* register-resident-weight kernel, folded layer 1 (H = 64 nets)  */
#define PSNODE_IMPL_TC 3       /* tcgen05 tensor-core kernel: 3xTF32, weights resident in TMEM (H = 64 ODE / DAE nets) */
#define PSNODE_IMPL_TC8 4      /* same, 8 warps per 16-trajectory group (4 accumulator elements per thread) */
#define PSNODE_IMPL_WIDE 5     /* tcgen05 kernels for the latent `*_02_direct_encode` nets (X = Z = H = 128, 2 layers): TMA-staged
                                  input series, hoisted input GEMM, both weight matrices resident in TMEM; also the 4-layer ODE_01 net
                                  at hidden <= 128 (X <= 16, Z <= 8; forward and tape-based reverse sweep, AUTO takes it for hidden 65..128) */

#define PSNODE_IMPL_LAYER 6    /* latent nets too wide for one SM (DAE_02 / ODE_02, X = Z (= V = I) = H = 128 or 256): one tcgen05 GEMM
                                  launch per layer over the whole batch shard, TMA-streamed operands, fused epilogues */

/* A small ELU MLP: Linear -> ELU -> ... -> Linear, weights in nn.Linear layout W[out][in] (row major,
 * contiguous), as built by the script-local DE_Func / AE_Func classes
 * (neural_00_ODE_01_no_encode.py:61-64, neural_00_ODE_02_direct_encode.py:52-53,
 *  neural_01_DAE_01_no_encode.py:64-67 and :77-80, neural_01_DAE_02_direct_encode.py:73-80 and :90-97). */
typedef struct psnode_mlp {
    int32_t n_layers;                      /* number of Linear layers, 1..PSNODE_MAX_LAYERS       */
    int32_t in_dim[PSNODE_MAX_LAYERS];
    int32_t out_dim[PSNODE_MAX_LAYERS];
    const float* W[PSNODE_MAX_LAYERS];
    const float* b[PSNODE_MAX_LAYERS];
} psnode_mlp;

/* strided (T, B, width) view, unit stride over the feature index */
typedef struct psnode_series {
    const float* p;
    int64_t st;     /* element stride between consecutive grid points */
    int64_t sb;     /* element stride between consecutive trajectories */
} psnode_series;

typedef struct psnode_series_out {
    float* p;
    int64_t st;
    int64_t sb;
} psnode_series_out;

/*
 * One integrate_ODE / integrate_DAE call.
 *   S = X + Z (+ V + I for a DAE) is the width of `all_initial` and of s = cat(x, held inputs).
 *   de : dx/dt network, input cat(a0, s - a0, s) (width 3S), output width X     (DE_Func.forward)
 *   ae : algebraic network, input cat(a0, x, z, v) (width S + X + Z + V), output width I (AE_Func.forward); DAE only
 *   ODE: initial state is x[0]; DAE: initial state is x_init, i_0 = ae(x_init, z[0], v[0]) (my_solvers.py:94-95).
 *   teacher_x / teacher_i reproduce input_true_x / input_true_i (my_solvers.py:72-74, 111-119, 121).
 *   event_idx[j] (j = 0..T-2) is the index k of the event that fires when LEAVING grid point j, or -1; it
 *   replaces the per-step host callback ODE_Event/DAE_Event.event_fn (neural_base.py:52-57, 180-185);
 *   the held inputs of that step are then z_jump[:,k] (and v_jump[:,k]) as in jump_change_fn (:59-65, :187-196).
 *   event_idx == NULL means "no event callbacks were given".
 */
typedef struct psnode_problem {
    int32_t kind;        /* PSNODE_ODE | PSNODE_DAE */
    int32_t method;      /* PSNODE_EULER | PSNODE_MIDPOINT | PSNODE_RK4 */
    int32_t impl;        /* PSNODE_IMPL_* */
    int32_t B;           /* trajectories */
    int32_t T;           /* grid points (N = T-1 steps) */
    int32_t X, Z, V, I;  /* widths; V = I = 0 for an ODE; Z may be 0 */
    int32_t teacher_x, teacher_i;
    int32_t E;           /* events per trajectory (second dim of z_jump / v_jump); 0 if none */
    psnode_series t;     /* (T,B,1) */
    psnode_series x;     /* (T,B,X): ODE initial state x[0]; teacher-forcing series when teacher_x */
    psnode_series z;     /* (T,B,Z) */
    psnode_series v;     /* (T,B,V)  DAE */
    psnode_series i;     /* (T,B,I)  DAE, read only when teacher_i */
    const float* x_init; /* (B,X)    DAE */
    int64_t x_init_sb;
    const float* a0;     /* (B,S) all_initial */
    int64_t a0_sb;
    const int32_t* event_idx;   /* [T-1] or NULL */
    const float* z_jump; /* (B,E,Z) */
    int64_t zj_sb, zj_se;
    const float* v_jump; /* (B,E,V) */
    int64_t vj_sb, vj_se;
    psnode_mlp de;
    psnode_mlp ae;
    psnode_series_out x_sol;    /* (T,B,X) */
    psnode_series_out i_sol;    /* (T,B,I)  DAE */
    /* Optional activation tape (caller-owned scratch, psnode_tape_floats(p) floats; NULL = none).  When given to
     * psnode_forward, the tensor-core kernel records the hidden activations of every RK stage -- what the reference's
     * autograd graph keeps alive between forward and loss.backward() (neural_00_ODE_01_no_encode.py:350-359) -- and
     * psnode_backward, given the SAME buffer, runs the tensor-core reverse sweep on it instead of recomputing the
     * stages from x_sol.  Results are identical up to fp32 rounding; without a tape nothing changes. */
    float* tape;
    int64_t tape_floats;
} psnode_problem;

/*
 * Reverse sweep (discrete adjoint = exact reverse mode of the unrolled loop; replaces the autograd graph
 * the reference builds and replays at loss.backward(), neural_00_ODE_01_no_encode.py:359 etc.).
 * `p` must describe the SAME problem as the forward call, with p->x_sol / p->i_sol holding the forward result
 * (they are the checkpoints the sweep restarts every step from).
 * Upstream gradients gx = dL/dx_sol (T,B,X) and gi = dL/di_sol (T,B,I) are strided views.
 * Outputs (any may be NULL = not wanted, except d_theta):
 *   d_theta   : flat fp32 vector, layout = for net in (de, ae): for layer: W (out*in, row major) then b (out)
 *   d_x0      : (B,X)  grad of x[0] (ODE) / x_init (DAE) as the INITIAL STATE
 *   d_a0      : (B,S)
 *   d_z, d_v  : (T,B,Z) / (T,B,V) grads of the held-input series (encoded variants: they carry grad)
 *   d_zjump, d_vjump : (B,E,Z) / (B,E,V)
 *   d_xteach, d_iteach : (T,B,X) / (T,B,I) grads of the teacher-forcing series (only with teacher_x / teacher_i)
 * All outputs are OVERWRITTEN (not accumulated).
 */
/* One masked squared-error term of the training scripts' loss (neural_00_ODE_01_no_encode.py:353-355,
 * neural_01_DAE_01_no_encode.py:414-418): sum_{j,b,c} w_c * mask[j,b] * (sol[j,b,c] - target[j,b,c])^2.
 * target = NULL pointer: term absent. */
typedef struct psnode_loss_term {
    psnode_series target;      /* (T,B,width) */
    psnode_series mask;        /* (T,B,1) */
    const float* feat_weight;  /* [width] or NULL = ones */
    const float* scale;        /* device scalar multiplying the term's gradient (upstream / sum(mask)); NULL = 1 */
} psnode_loss_term;

typedef struct psnode_adjoint {
    psnode_series gx;
    psnode_series gi;
    float* d_theta;
    int64_t n_theta;
    float* d_x0;      int64_t d_x0_sb;
    float* d_a0;      int64_t d_a0_sb;
    psnode_series_out d_z;
    psnode_series_out d_v;
    float* d_zjump;   int64_t d_zj_sb, d_zj_se;
    float* d_vjump;   int64_t d_vj_sb, d_vj_se;
    psnode_series_out d_xteach;
    psnode_series_out d_iteach;
    /* Loss fusion: when fuse_x.target.p != NULL the sweep IGNORES gx and forms the upstream gradient
     *     dL/dx_sol[j,b,c] = scale[0] * 2 * w_c * mask[j,b] * (x_sol[j,b,c] - target[j,b,c])
     * on the fly from p->x_sol, so the (T,B,X) gradient tensor of the masked MSE is never materialised; fuse_i likewise
     * replaces gi (DAE).  Supported by the tensor-core sweeps (psnode_sweep_fuses_loss tells); otherwise pass gx / gi. */
    psnode_loss_term fuse_x;
    psnode_loss_term fuse_i;
} psnode_adjoint;

/* library / device introspection */
int psnode_abi_version(void);
const char* psnode_status_string(int status);
const char* psnode_last_cuda_error(void);
/* number of kernel launches issued by this library since load (bench.py's gpu_launches counter) */
int64_t psnode_launch_count(void);
/* name of the kernel variant the last psnode_forward / psnode_backward call dispatched to */
const char* psnode_last_kernel(void);

/* number of parameters of a net = sum(out*in + out) */
int64_t psnode_mlp_param_count(const psnode_mlp* m);

/*
 * Build the per-step event table on the device (no host sync).
 *   t0      : times of SAMPLE 0, element j at t0[j*t_st]          (reference looks at sample 0 only: neural_base.py:54)
 *   ev0     : event times of SAMPLE 0, element k at ev0[k*ev_se]  (event_t[0], shape (E,1))
 *   event_idx[j] = k if t0[j] == ev0[k] (exact float equality) else -1, for j = 0..T-2
 *   err[0] is set to 1 if some grid point matches more than one event (the reference raises there), else 0.
 * Replaces ODE_Event.event_fn / DAE_Event.event_fn being called from the hot loop every step (my_solvers.py:70, :108).
 */
int psnode_event_table(const float* t0, int64_t t_st, int32_t T, const float* ev0, int64_t ev_se, int32_t E,
                       int32_t* event_idx, int32_t* err, void* stream);

/* floats of activation tape psnode_forward can record for this problem (0: this problem has no tape-based reverse
 * sweep, psnode_backward recomputes from x_sol) */
int64_t psnode_tape_floats(const psnode_problem* p);

/* 1 if the tape-based reverse sweep of this problem also produces the input-series / jump gradients (d_z, d_zjump) --
 * the latent `*_02_direct_encode` nets, whose encoders always need them (neural_00_ODE_02_direct_encode.py:75-86);
 * 0: requesting those gradients sends psnode_backward to the recomputing sweep and a tape would be wasted. */
int psnode_tape_covers_input_grads(const psnode_problem* p);

/* 1 if psnode_backward(p, a) will run a sweep that honours a->fuse_x / a->fuse_i (the tape-based tensor-core sweeps) */
int psnode_sweep_fuses_loss(const psnode_problem* p, const psnode_adjoint* a);

/* workspace sizes in bytes (0 is possible) */
int64_t psnode_forward_workspace(const psnode_problem* p);
int64_t psnode_backward_workspace(const psnode_problem* p, const psnode_adjoint* a);

/* forward integration: FixedGridODESolver.integrate_ODE / integrate_DAE (my_solvers.py:52-80, 82-131) incl. the
 * per-step step_integrate (:48-50) -> _step_func (my_fixed_grid.py) -> DE_Func/AE_Func forward evaluations. */
int psnode_forward(const psnode_problem* p, void* workspace, int64_t workspace_bytes, void* stream);

/* reverse sweep, see psnode_adjoint */
int psnode_backward(const psnode_problem* p, const psnode_adjoint* a, void* workspace, int64_t workspace_bytes,
                    void* stream);

/*
 * Host-buffer convenience entry (the end-to-end path bench.py times as `e2e`): every data pointer in `p`
 * (series, x_init, a0, jumps, weights, outputs) is a HOST pointer, event_idx is a HOST table or NULL.  Inputs
 * are staged to the device with cudaMemcpyAsync on `stream`, the problem is integrated, and x_sol / i_sol are copied
 * back; the call returns after the stream has drained.  `h2d_bytes` / `d2h_bytes` receive the bytes moved.
 * This is the only entry point that allocates (device scratch, cached across calls) and synchronises.
 */
int psnode_forward_host(const psnode_problem* p, void* stream, int64_t* h2d_bytes, int64_t* d2h_bytes);

/*
 * Masked squared-error loss numerator of the training scripts and its gradient, one pass over the trajectory each
 * (SURVEY 8f next-2).  Replaces `torch.sum(Loss_func(x_pred, x, reduction='none') * mask)` and its autograd backward
 * (neural_00_ODE_01_no_encode.py:353-355; neural_01_DAE_01_no_encode.py:414-418 with per-feature weights):
 *     loss[0] = sum_{o,i,c} w_c * mask[o,i] * (pred[o,i,c] - target[o,i,c])^2          (w = NULL: all ones)
 *     grad[o,i,c] = upstream[0] * 2 * w_c * mask[o,i] * (pred[o,i,c] - target[o,i,c])
 * The three inputs are strided (n_outer, n_inner, X) views with unit feature stride (`st` = outer stride, `sb` = inner
 * stride; mask has width 1); pass the dimension with the smaller stride as the inner one -- (T,B,X) solver output or the
 * scripts' batch-major (B,T,X) both work.  `loss`, `upstream` are device scalars (no host synchronisation); the caller
 * divides by sum(mask).  Deterministic summation order.
 */
int64_t psnode_masked_sse_workspace(void);
int psnode_masked_sse(const psnode_series* pred, const psnode_series* target, const psnode_series* mask,
                      const float* feat_weight, int32_t n_outer, int32_t n_inner, int32_t X, float* loss,
                      void* workspace, int64_t workspace_bytes, void* stream);
int psnode_masked_sse_grad(const psnode_series* pred, const psnode_series* target, const psnode_series* mask,
                           const float* feat_weight, int32_t n_outer, int32_t n_inner, int32_t X,
                           const float* upstream, const psnode_series_out* grad, void* stream);

/*
 * Encoder / decoder fusion for the `*_02_direct_encode` models (ODE_Model.forward,
 * neural_00_ODE_02_direct_encode.py:75-89; DAE_Model.forward, neural_01_DAE_02_direct_encode.py:126-153).  The scripts encode the raw
 * input series with 2-layer ELU MLPs into (T,B,H) latent tensors, call integrate_ODE / integrate_DAE on them and decode the latent
 * trajectory with 2-layer MLPs.  psnode_forward_encoded does the three steps in one call, time chunk by time chunk, so that no
 * (T,B,H) tensor ever exists in HBM:
 *   - z_enc / v_enc: Linear(raw -> H) . ELU . Linear(H -> H).  The second Linear is folded into the held-input half of layer 1
 *     (F_z E2) and the hidden layer ELU(E1 raw + e1) is generated in shared memory as the GEMM operand from the raw
 *     (T,B,<=8) series;
 *   - the integration runs on `chunk_rows` grid rows at a time (latent rows live in a chunk-sized scratch);
 *   - x_dec / i_dec: Linear(H -> H) . ELU . Linear(H -> raw width <= 128) applied to every latent row before the store.
 * In `p`: z.p / v.p / z_jump / v_jump / x_sol.p / i_sol.p are ignored (NULL); Z = V = I = X = H in {128, 256}; x_init (DAE) or x.p
 * row 0 (ODE) is the LATENT initial state and a0 the latent all_initial (the caller encodes those B rows); event_idx as usual.
 * Forward / evaluation only (no reverse sweep through the fused codecs: training uses the unfused calls).
 */
typedef struct psnode_codec {
    int32_t ZR, VR, XR, IR;            /* raw widths: inputs z_raw, v_raw (1..8); decoded outputs (1..128); VR = IR = 0 for an ODE */
    psnode_series z_raw, v_raw;        /* (T,B,ZR), (T,B,VR) */
    const float* zj_raw;               /* (B,E,ZR) raw jump values (events) */
    int64_t zjr_sb, zjr_se;
    const float* vj_raw;               /* (B,E,VR) */
    int64_t vjr_sb, vjr_se;
    psnode_mlp z_enc, v_enc;           /* 2 layers each */
    psnode_mlp x_dec, i_dec;           /* 2 layers each */
    psnode_series_out x_out, i_out;    /* decoded trajectories (T,B,XR), (T,B,IR) */
    int32_t chunk_rows;                /* grid rows per time chunk; 0 = default (64) */
} psnode_codec;

int64_t psnode_forward_encoded_workspace(const psnode_problem* p, const psnode_codec* c);
int psnode_forward_encoded(const psnode_problem* p, const psnode_codec* c, void* workspace, int64_t workspace_bytes, void* stream);

/*
 * Init_Func + all_initial construction in one launch (DAE_Model.forward, neural_01_DAE_01_no_encode.py:50-58, :98-99;
 * neural_01_DAE_02_direct_encode.py:126-127):
 *     x0 = init(cat(z0, v0, i0))   (Linear/ELU chain),     all_initial = cat(x0, z0, v0, i0)
 * z0 / v0 / i0 are the first grid rows of the series, (B, width) with row stride *_sb (unit feature stride).
 * psnode_init_state_backward is the exact reverse mode: d_x0 (B,X) = gradient of the initial state (psnode_adjoint.d_x0) and d_a0 (B,S)
 * = gradient of all_initial (either may be NULL) -> d_theta (Init_Func parameters, layout as psnode_adjoint.d_theta for one net) and
 * the gradients of the three rows (each may be NULL).  Deterministic.
 */
int psnode_init_state(const psnode_mlp* init, const float* z0, int64_t z_sb, const float* v0, int64_t v_sb, const float* i0, int64_t i_sb,
                      int32_t B, int32_t Z, int32_t V, int32_t I, float* x0, int64_t x0_sb, float* a0, int64_t a0_sb, void* stream);
int64_t psnode_init_state_backward_workspace(const psnode_mlp* init, int32_t B);
int psnode_init_state_backward(const psnode_mlp* init, const float* z0, int64_t z_sb, const float* v0, int64_t v_sb, const float* i0, int64_t i_sb,
                               int32_t B, int32_t Z, int32_t V, int32_t I, const float* d_x0, int64_t d_x0_sb, const float* d_a0, int64_t d_a0_sb,
                               float* d_theta, float* d_z0, int64_t d_z0_sb, float* d_v0, int64_t d_v0_sb, float* d_i0, int64_t d_i0_sb,
                               void* workspace, int64_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PSNODE_B200_H */
