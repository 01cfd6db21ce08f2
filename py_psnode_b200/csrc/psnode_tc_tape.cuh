// psnode_tc_tape.cuh -- the activation tape shared by the tensor-core forward kernel (psnode_tc_fwd.cu, writer) and the
// tensor-core reverse sweep (psnode_tc_bwd.cu, reader).
//
// The reference back-propagates through its unrolled Python loop, i.e. autograd keeps EVERY intermediate of every stage of
// every step alive (neural_00_ODE_01_no_encode.py:359 `loss.backward()`).  The generic reverse sweep here recomputes them
// from the stored trajectory instead; on a B200 (180 GB HBM3e, 6.5 TB/s) the cheaper choice for the H = 64 nets is the
// reference's: the forward kernel records the three post-ELU hidden activations and the stage input of every RK stage
// (3328 floats per stage and 16-trajectory group = 3.3 KB per trajectory-step, 13.6 GB at B = 4096 x 1000 RK4 steps), and the
// reverse sweep streams them back -- 27 GB of extra HBM traffic per training step (4 ms at the measured copy bandwidth)
// instead of a second forward integration (10.7 ms) plus the shared-memory footprint of a second weight set.
//
// Layout: every value is written and later read by the SAME thread position of the 128-thread group (thread-private, fully
// coalesced 32-byte / 8-byte accesses):
//   record(group gid, step j = 1..T-1, stage e) at float offset ((gid * (T-1) + (j-1)) * NST + e) * PSN_TAPE_STAGE
//     [   0, 1024) a1 fragments: thread gt holds 8 floats at gt*8   (element i <-> row m0 + 8*((i>>1)&1), trajectory
//     [1024, 2048) a2 fragments                                       c0 + (i&1) + 8*(i>>2); m0 = 16*warp + lane/4, c0 = 2*(lane%4))
//     [2048, 3072) a3 fragments
//     [3072, 3328) stage input y_e: thread gt holds states (lane/4, lane/4 + 8) of trajectory c0 + (warp&1) + 8*(warp>>1) at gt*2
#pragma once
#include <stdint.h>

constexpr int PSN_TAPE_FRAG = 1024;                       // floats per hidden layer and group
constexpr int PSN_TAPE_STAGE = 3 * PSN_TAPE_FRAG + 256;   // floats per (group, step, stage)
constexpr int PSN_TC_TN = 16;                             // trajectories per group

// number of 16-trajectory groups and how many of them share a CTA (identical in the forward and the reverse kernel)
static inline int psn_tc_ngroups(int B) { return (B + PSN_TC_TN - 1) / PSN_TC_TN; }
static inline int psn_tc_groups_per_cta(int B) { return psn_tc_ngroups(B) > 148 ? 2 : 1; }
static inline int psn_tc_nstages(int method) { return method == 0 ? 1 : (method == 1 ? 2 : 4); }
static inline int64_t psn_tc_tape_floats(int B, int T, int method) {
    return (int64_t)psn_tc_ngroups(B) * (T > 1 ? T - 1 : 0) * psn_tc_nstages(method) * PSN_TAPE_STAGE;
}

// ---- DAE tape (integrate_DAE, my_solvers.py:82-131): the same PSN_TAPE_STAGE-sized records, per 16-trajectory group ----
//   step j (1..T-1): NST stage records, then the record of the algebraic evaluation i_j = ae(x_j, z[j], v[j])   (:121)
//   then one record for i_0 = ae(x_0, z[0], v[0]) (:95) and one per event k for the re-evaluated i_0 (:108-110), whose
//   "stage input" slot holds that i_0 (thread (w, h, lane) owns row lane/4 + 8h of trajectory c0 + (w&1) + 8(w>>1)).
__host__ __device__ static inline int psn_dae_recs_per_step(int nst) { return nst + 1; }
__host__ __device__ static inline int64_t psn_dae_group_recs(int T, int nst, int E) {
    return (int64_t)(T > 1 ? T - 1 : 0) * (nst + 1) + 1 + (E > 0 ? E : 0);
}
__host__ __device__ static inline int64_t psn_dae_rec_stage(int j, int e, int nst) { return (int64_t)(j - 1) * (nst + 1) + e; }
__host__ __device__ static inline int64_t psn_dae_rec_point(int j, int T, int nst) {       // j = 0..T-1
    return j == 0 ? (int64_t)(T > 1 ? T - 1 : 0) * (nst + 1) : (int64_t)(j - 1) * (nst + 1) + nst;
}
__host__ __device__ static inline int64_t psn_dae_rec_event(int k, int T, int nst) {
    return (int64_t)(T > 1 ? T - 1 : 0) * (nst + 1) + 1 + k;
}
static inline int64_t psn_tc_dae_tape_floats(int B, int T, int method, int E) {
    return (int64_t)psn_tc_ngroups(B) * psn_dae_group_recs(T, psn_tc_nstages(method), E) * PSN_TAPE_STAGE;
}
