/*
 * liblime_b200 -- C ABI of the B200-native (sm_100a) density-matrix engine that replaces
 * the CPU hot path of binggu56/lime.
 *
 * lime has no FFI of its own: its boundary is the Python API.  Each entry point below
 * names the reference function whose inner loop it replaces (file:line in the reference
 * tree); the modules of lime_b200/ re-creates those Python signatures and calls into this library
 * through ctypes (see INTEGRATION.md for the binding a lime maintainer would add).
 *
 * Conventions
 *   - complex128 data are interleaved (re,im) doubles, row-major -- the memory layout of
 *     a C-contiguous numpy complex128 array; pointers named h_* are HOST pointers, d_*
 *     are DEVICE pointers on the plan's device; `stream` is a cudaStream_t (0 = default).
 *   - every function returns 0 on success, <0 on error (limeb200_last_error() has the
 *     text, thread-local); nothing throws; launches are asynchronous on `stream`.
 *   - the caller owns every buffer it passes; a plan owns only its operator copies and
 *     scratch.  Plans may be used concurrently from different streams only if distinct.
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef LIME_B200_H
#define LIME_B200_H
#ifdef __cplusplus
extern "C" {
#endif

#define LIMEB200_VERSION 100

int limeb200_version(void);
const char* limeb200_last_error(void);
/* sm count, compute capability and opt-in shared memory per block of `device` */
int limeb200_device_info(int device, int* sm_count, int* cc_major, int* cc_minor,
                         long long* smem_optin, long long* l2_bytes);

/* ------------------------------------------------------------------------------------
 * Quantum master equation in generator/sandwich form
 *
 *      d rho/dt = G rho + rho Gr + sum_s X_s rho Z_s^H        (Gr = G^H unless set)
 *
 * replaces: liouvillian/lindbladian  lime/oqs.py:706-723 (= lime/phys.py:561-577)
 *             G = -iH - 1/2 sum_m l_m^H l_m,   (X_s, Z_s) = (l_s, l_s)
 *           func (Redfield, operator form)     lime/oqs.py:840-850; same generator as the
 *           N^2 x N^2 tensor of redfield_tensor lime/oqs.py:528-579:
 *             G = -i diag(eps) - sum_k A_k Lam_k, sandwiches (A_k, Lam_k), (Lam_k, A_k)
 *           rk4                                  lime/phys.py:636-649
 *           the time loops of _lindblad          lime/oqs.py:1674-1682
 *                             _lindblad_driven   lime/oqs.py:1782-1795
 *                             _redfield          lime/oqs.py:453-462
 *           obs_dm                               lime/phys.py:837-844
 * ------------------------------------------------------------------------------------ */
typedef struct limeb200_qme_s* limeb200_qme_t;

int limeb200_qme_create(limeb200_qme_t* plan, int N, int device);
int limeb200_qme_destroy(limeb200_qme_t plan);

/* Operators are copied at the call.  `nb` = 1 (operator shared by every density matrix of
 * the batch) or the batch size B (one set of VALUES per density matrix, same sparsity).
 * dense : h_* is [nb][N][N] complex.
 * csr   : scipy layout, pattern shared by the batch: indptr[N+1], indices[nnz],
 *         h_data [nb][nnz] complex.                                                      */
int limeb200_qme_set_generator_dense(limeb200_qme_t plan, const double* h_G, int nb);
int limeb200_qme_set_generator_csr(limeb200_qme_t plan, const int* indptr, const int* indices,
                                   const double* h_data, int nnz, int nb);
int limeb200_qme_add_sandwich_dense(limeb200_qme_t plan, const double* h_X, const double* h_Z, int nb);
int limeb200_qme_add_sandwich_csr(limeb200_qme_t plan,
                                  const int* x_indptr, const int* x_indices, const double* h_xdata, int x_nnz,
                                  const int* z_indptr, const int* z_indices, const double* h_zdata, int z_nnz,
                                  int nb);
/* optional explicit right generator (dense paths only).  lime's liouvillian is the plain
 * commutator -i(H rho - rho H) even for a non-Hermitian H (lime/oqs.py:706-713), i.e.
 * Gr = +iH - 1/2 sum l^H l, which equals G^H only when H is Hermitian.                    */
int limeb200_qme_set_right_generator_dense(limeb200_qme_t plan, const double* h_Gr, int nb);
/* time-dependent generator, frozen over the four stages of step k (lime/oqs.py:1786-1790
 * evaluates H(t) once per step at t+dt):  G_k = G + sum_i coef[k][i] D_i and
 *   h_Dr given : Gr_k = Gr + sum_i coef[k][i] Dr_i      (complex drives, lime's Pulse.efield)
 *   h_Dr NULL  : Gr_k = Gr + sum_i conj(coef[k][i]) D_i^H
 * D_i, Dr_i dense [N][N].                                                                  */
int limeb200_qme_add_drive_dense(limeb200_qme_t plan, const double* h_D, const double* h_Dr);
/* observables e[E][N][N] (dense host): obs = Tr(e rho), lime/phys.py:837-844 */
int limeb200_qme_set_observables(limeb200_qme_t plan, const double* h_e, int E);
/* kernel selection: 0 auto, 1 dense on-chip, 2 dense stage-wise, 3 sparse global-scratch,
 * 4 sparse cluster-resident (generic ELL), 5 sparse cluster-resident register-tiled
 * (at most 4 off-diagonal entries per row of G, one entry per row of X_s/Z_s, N <= 128);
 * tests force each path                                                                  */
/* The generator (and sandwich) batch index is the RK4 STEP instead of the density matrix: values [nsteps][nnz] of a
 * time-dependent sparse generator, frozen over the four stages of a step -- _lindblad_driven with CSR operands,
 * lime/oqs.py:1691-1800 (Hermitian H(t): real drive envelopes).  Selects the sparse global-scratch kernel.    */
int limeb200_qme_set_step_values(limeb200_qme_t plan, int on);
int limeb200_qme_set_path(limeb200_qme_t plan, int path);
/* analyse operators, choose the kernel, upload.  B_hint sizes scratch (may grow later). */
int limeb200_qme_finalize(limeb200_qme_t plan);
/* which kernel finalize chose (same numbering as set_path) */
int limeb200_qme_get_path(limeb200_qme_t plan);
/* what finalize decided: info9 = {path, basis permuted?, bandwidth (halo rows), band kernel: off-diagonal slots,
 * purely imaginary off-diagonal G?, real X/Z?, cluster size, rows per CTA, chain stride}; perm[N] = new -> old basis
 * order (may be NULL).  limeb200_qme_create(.., device = -1) makes an ANALYSIS-ONLY plan (no GPU needed, assumes a
 * B200: 148 SMs, 227 KB shared memory) on which this is the only useful call -- it cannot run.          */
int limeb200_qme_get_info(limeb200_qme_t plan, int* info9, int* perm);

/* nsteps RK4 steps of B density matrices, in place in d_rho[B][N][N].
 *   d_coef : [nsteps][ndrive] complex or NULL
 *   d_obs  : [nsteps][B][E] complex or NULL; sample k is the state AFTER step k+1
 *   d_traj : [nsteps/traj_every][B][N][N] or NULL (rholist of lime's Result)               */
int limeb200_qme_run(limeb200_qme_t plan, double* d_rho, int B, double dt, int nsteps,
                     const double* d_coef, double* d_obs, double* d_traj, int traj_every,
                     void* stream);
/* one right-hand side: d_out[B][N][N] = L(d_in) -- lime/oqs.py:706-713 */
int limeb200_qme_rhs(limeb200_qme_t plan, const double* d_in, double* d_out, int B, void* stream);
/* number of kernels the last run launched (bench.py's gpu_launches) */
long long limeb200_qme_last_launches(limeb200_qme_t plan);

/* plain batched complex GEMM on the FP64 tensor cores (DMMA): C[b] = A[b] * B[b], row-major,
 * A [M][K], B [K][N], C [M][N]; strides in complex elements, 0 = shared by the batch.
 * replaces the dense products of the Liouvillian eigen-decomposition solver,
 * lime/superoperator.py:525-560 (evolve) and :703-754 (correlation_3op_2t: tmp1.T @ coeff @ tmp2) */
int limeb200_zgemm(const double* d_A, const double* d_B, double* d_C, int M, int N, int K, int batch,
                   long long sA, long long sB, long long sC, void* stream);

/* ------------------------------------------------------------------------------------
 * Runge-Kutta-Fehlberg 4(5): fused stage updates on device-resident real vectors of n doubles
 * (a complex128 state is 2 x its element count).
 * replaces: the stage arithmetic of r8_rkf45 / r8_fehl, the integrator lime's examples/rkf45_test.py:7,115
 *           drives (`from lime.rkf45 import *`; the module itself is not in the lime tree).  Step-size control
 *           is host logic (lime_b200/rkf45.py); the right-hand side is any device function, e.g.
 *           limeb200_qme_rhs / limeb200_heom_rhs.
 *   stage 1..5 : d_out = y + h * (Fehlberg combination of yp, f1..f(stage-1)) -- the argument of slope `stage`
 *                (stage 5 yields the argument of the sixth function evaluation); unused slopes may be NULL
 *   error      : d_s = 5th-order-pair solution; d_result2[0] = max_k ee_k/et_k, d_result2[1] = min_k et_k with
 *                et = |y| + |s| + ae, ee = |-2090 yp + 21970 f3 - 15048 f4 + 22528 f2 - 27360 f5|
 *   hinit      : d_result2[0] = max_k tol_k, d_result2[1] = min(h0, min_k (tol_k/|yp_k|)^(1/5) where tol_k < |yp_k| h0^5),
 *                tol_k = relerr |y_k| + abserr  (the start-up step size rule)
 *   axpy       : y += a x
 * ------------------------------------------------------------------------------------ */
int limeb200_rkf45_stage(int stage, long long n, const double* d_y, const double* d_yp, const double* d_f1,
                         const double* d_f2, const double* d_f3, const double* d_f4, const double* d_f5, double h,
                         double* d_out, void* stream);
int limeb200_rkf45_error(long long n, const double* d_y, const double* d_yp, const double* d_f2, const double* d_f3,
                         const double* d_f4, const double* d_f5, double h, double ae, double* d_s,
                         double* d_result2, void* stream);
int limeb200_rkf45_hinit(long long n, const double* d_y, const double* d_yp, double relerr, double abserr, double h0,
                         double* d_result2, void* stream);
int limeb200_rkf45_axpy(long long n, double a, const double* d_x, double* d_y, void* stream);

/* ------------------------------------------------------------------------------------
 * Liouville-space linear ODE  dv/dt = R v,  R an arbitrary D x D CSR matrix (DEVICE arrays)
 * replaces: rhs + rk4 loop of _redfield  lime/oqs.py:453-472 (R from redfield_tensor),
 *           expm(method='EOM')            lime/phys.py:1384-1401 (B = D unit vectors)
 *   d_v   : [B][D] in/out
 *   d_e   : [E][D] complex row vectors, obs = sum_a e[a] v[a]  (no conjugation)
 *   d_obs : [nsteps][B][E];  d_traj: [nsteps/traj_every][B][D]
 * ------------------------------------------------------------------------------------ */
int limeb200_liouville_rk4_csr(const int* d_indptr, const int* d_indices, const double* d_data,
                               int D, double* d_v, int B, const double* d_e, int E,
                               double* d_obs, double* d_traj, int traj_every,
                               double dt, int nsteps, void* stream);

/* ------------------------------------------------------------------------------------
 * HEOM
 * ------------------------------------------------------------------------------------ */
/* Index tables -- HOST functions, bit-exact with state_number_enumerate /
 * enr_state_dictionaries, lime/heom/heom.py:21-108 (lexicographic, last index fastest,
 * prefix pruning; excitations == 0 means unrestricted, lime/heom/heom.py:61).
 * count first, then fill states[nhe][nmodes], dn[nhe][nmodes], up[nhe][nmodes]
 * (dn = index of n - e_k or -1; up = index of n + e_k or -1; connectivity rules
 * lime/heom/heom.py:176-216).                                                            */
long long limeb200_heom_count_states(const int* dims, int nmodes, int excitations);
int limeb200_heom_build_tables(const int* dims, int nmodes, int excitations, long long nhe,
                               int* states, int* dn, int* up);

typedef struct limeb200_heom_s* limeb200_heom_t;
/* Multi-index hierarchy (rules lime/heom/heom.py:156-216 + system term lime/oqs.py:1854):
 *   d rho_n/dt = -i[H,rho_n] - (sum_k n_k nu_k) rho_n
 *              + sum_k pref_dn n_k (c_k Q_k rho_{n-e_k} - conj(c_k) rho_{n-e_k} Q_k)
 *              + sum_k pref_up [Q_k, rho_{n+e_k}],        Q_k = Q[qmap[k]]
 * h_H [n][n], h_Q [nq][n][n], h_c [nmodes] complex, h_nu [nmodes] real,
 * pref_dn / pref_up complex (re,im) -- (-i,-i) for the rules as written in lime.
 * Tables as produced by limeb200_heom_build_tables (any consistent tables are accepted).
 * `row_lo,row_hi`: the ADO range this plan OWNS (0,nhe for a single GPU); see
 * limeb200_heom_stage for the sharded protocol.                                          */
int limeb200_heom_create(limeb200_heom_t* plan, int device, int n, int nmodes, int nq, long long nhe,
                         const double* h_H, const double* h_Q, const int* qmap,
                         const double* h_c, const double* h_nu,
                         const double* pref_dn, const double* pref_up,
                         const int* states, const int* dn, const int* up,
                         long long row_lo, long long row_hi);
/* same, with per-hierarchy bath parameters: h_c [npar][nmodes], h_nu [npar][nmodes],
 * npar = 1 or the batch size B passed to limeb200_heom_run (parameter sweeps)             */
int limeb200_heom_create_batched(limeb200_heom_t* plan, int device, int n, int nmodes, int nq, long long nhe,
                                 const double* h_H, const double* h_Q, const int* qmap,
                                 const double* h_c, const double* h_nu, int npar,
                                 const double* pref_dn, const double* pref_up,
                                 const int* states, const int* dn, const int* up,
                                 long long row_lo, long long row_hi);
int limeb200_heom_destroy(limeb200_heom_t plan);
/* 0 auto, 1 on-chip (one CTA per hierarchy, all steps fused), 2 one launch per RK4 stage,
 * 3 persistent cooperative kernel (all steps in one launch, one grid barrier per stage),
 * 4 dataflow-synchronised persistent kernel (one hierarchy, diagonal coupling operators: tagged stage
 *   vectors, no barrier; opt-in on one GPU, the default of the ADO-sharded multi-GPU propagator)   */
int limeb200_heom_set_path(limeb200_heom_t plan, int path);
int limeb200_heom_get_path(limeb200_heom_t plan);
/* nsteps RK4 steps (lime/phys.py:636-649) of B hierarchies d_ado[B][nhe][n][n], in place.
 * d_eT [E][n][n] TRANSPOSED observables of tier 0, d_obs [nsteps][B][E],
 * d_traj [nsteps/traj_every][B][n][n] tier-0 trajectory.                                  */
int limeb200_heom_run(limeb200_heom_t plan, double* d_ado, int B, double dt, int nsteps,
                      const double* d_eT, int E, double* d_obs, double* d_traj, int traj_every,
                      void* stream);
/* one right-hand side over the owned range: d_out[B][nhe][n][n] (rows outside the range untouched) */
int limeb200_heom_rhs(limeb200_heom_t plan, const double* d_in, double* d_out, int B, void* stream);
/* one RK4 stage over the owned ADO range (building block of the ADO-sharded propagator:
 * the host exchanges d_ynext between ranks after each call).  stage = 0..3.
 *   stage 0 reads d_rho as the stage vector; others read d_yin.                           */
int limeb200_heom_stage(limeb200_heom_t plan, int stage, double* d_rho, const double* d_yin,
                        double* d_ynext, double* d_acc, int B, double dt, void* stream);
long long limeb200_heom_last_launches(limeb200_heom_t plan);

/* ADO-sharded propagation of ONE hierarchy over the GPUs of one box (one process per GPU): the
 * fused compute + exchange path.  Every rank runs a persistent kernel over the ADO range its plan
 * owns; each new stage-vector element is stored into the local stage vector AND, through peer
 * (CUDA IPC / NVLink) pointers, into the same element of every peer's stage vector; the RK4 stages are
 * separated by a one-hop barrier: every CTA counts itself on every peer with one remote reduction and waits, in its
 * OWN memory, for all CTAs of all ranks.  No host round trip and no collective library
 * call inside the time loop (the alternative is limeb200_heom_stage + an all-gather per stage).
 *   limeb200_peer_alloc : cudaMalloc + zero + export (handle64 = cudaIpcMemHandle_t bytes)
 *   limeb200_peer_open  : map a peer's allocation into this process
 *   d_y0/d_y1[world]    : every rank's two stage vectors [nhe_pad][n][n] (index = rank; own entry local)
 *   d_flags[world]      : every rank's flag array (>= world unsigned, zero-initialised): word q counts arrivals of rank q
 *   h_grids[world]      : HOST array, limeb200_heom_persist_grid(plan_r, 1) of every rank r (arrivals per stage)
 *   d_rho               : local [nhe_pad][n][n]; only the owned rows are read and written
 *   d_peer_mask         : [nhe] bytes or NULL.  Bit q of entry a = "peer slot q (the q-th rank != this one) reads
 *                         ADO a" (it owns a neighbour of a): new values of a are stored only into those peers, except
 *                         at the last stage of the run, which goes to every peer.  NULL: always every peer.
 *   epoch               : flags are monotonic; pass the number of stages run so far (4 * steps)
 * On entry every rank's d_y0[rank] holds the full state; on return it holds the full new state.  */
int limeb200_memcpy_d2d(void* d_dst, const void* d_src, long long bytes, void* stream);   /* async, stream-ordered */
int limeb200_peer_alloc(int device, long long bytes, void** d_ptr, unsigned char* handle64);
int limeb200_peer_open(int device, const unsigned char* handle64, void** d_ptr);
int limeb200_peer_close(int device, void* d_ptr);
int limeb200_peer_free(int device, void* d_ptr);
int limeb200_heom_persist_grid(limeb200_heom_t plan, int B);   /* CTAs the persistent kernel uses for B hierarchies (>= 1), <0 = error */
int limeb200_heom_run_sharded(limeb200_heom_t plan, int rank, int world, void* const* d_y0, void* const* d_y1,
                              void* const* d_flags, const int* h_grids, double* d_rho, const unsigned char* d_peer_mask,
                              double dt, int nsteps, unsigned epoch, void* stream);
/* Dataflow variant of the sharded propagator (csrc/heom_flow.cuh): no barrier between the stages.  Stage vectors
 * are TAGGED -- entry e of d_T0 / d_T1 is two 16-byte words {value bits, 64-bit stage tag} -- and a consumer polls, in
 * its own memory, exactly the entries it reads; producers store new entries locally and into the buffers of the ranks
 * that read them (d_peer_mask as above; the last stage of a run goes to every rank).
 *   d_T0 / d_T1[world] : every rank's two tagged buffers, 32 * nhe * n * n bytes each, zero-initialised (peer_alloc)
 *   flow_pack          : d_T0 <- full state d_y [nhe][n][n] tagged `tag` (every rank, BEFORE the ranks synchronise
 *                        on the host and launch flow_run_sharded with tag0 = tag)
 *   flow_run_sharded   : nsteps RK4 steps; d_rho [nhe][n][n] local, owned rows read and written; tags used are
 *                        tag0 .. tag0 + 4 nsteps: the next run must use a larger tag0
 *   flow_unpack        : d_y <- values of d_T0 (after the ranks have synchronised on the host), checking every tag
 *   limeb200_heom_sharded_error reports bit 0 = a wait timed out, bit 1 = unpack met a stale tag               */
int limeb200_heom_flow_supported(limeb200_heom_t plan);   /* 0 no, 1 tiled variant, 2 register-resident variant */
int limeb200_heom_flow_pack(limeb200_heom_t plan, const double* d_y, void* d_T0, unsigned long long tag, void* stream);
int limeb200_heom_flow_unpack(limeb200_heom_t plan, const void* d_T0, unsigned long long tag, double* d_y, void* stream);
int limeb200_heom_flow_run_sharded(limeb200_heom_t plan, int rank, int world, void* const* d_T0, void* const* d_T1,
                                   double* d_rho, const unsigned char* d_peer_mask, double dt, int nsteps,
                                   unsigned long long tag0, void* stream);
/* Hybrid of the two: the barrier kernel (plain 16-byte elements, grid barrier) INSIDE each GPU, the tagged halo of the
 * dataflow kernel BETWEEN the GPUs -- rows owned by other ranks are polled in the tagged inbox d_T0 / d_T1 (buffers and
 * tags as for flow_run_sharded; use flow_pack before and flow_unpack after), no cross-GPU barrier.  For shards that
 * are too large for the dataflow kernel's one-element-per-thread regime.                                             */
int limeb200_heom_run_sharded_halo(limeb200_heom_t plan, int rank, int world, void* const* d_T0, void* const* d_T1,
                                   double* d_rho, const unsigned char* d_peer_mask, double dt, int nsteps,
                                   unsigned long long tag0, void* stream);
/* 1 when a bounded spin of the last sharded run timed out (a peer never arrived), else 0 */
int limeb200_heom_sharded_error(limeb200_heom_t plan, void* stream);

/* _heom_dl-exact propagator, lime/oqs.py:1802-1865: single Drude mode, in-place
 * Gauss-Seidel Euler sweep over a linear chain of `nado` tiers (tier 0 advanced twice per
 * step, last tier frozen).  d_ado[B][nado][n][n] (tier-major, unlike lime's (n,n,tier)),
 * h_H[n][n], h_sz[n][n]; per-hierarchy parameters d_par[B][3] = (gamma, a, b) with
 * a = pi*lambda*T, b = 0 in lime.  d_traj[nt][B][n][n] = tier 0 after every step or NULL. */
int limeb200_heom_dl_euler(const double* h_H, const double* h_sz, int n, int nado,
                           double* d_ado, const double* d_par, int B, double dt, int nt,
                           double* d_traj, void* stream);

/* ------------------------------------------------------------------------------------
 * Sum-over-states response functions (lime/signal/sos.py:230-283, 348-729, 904-1099)
 * in factorised form.  Every third-order pathway of sos.py is
 *      S_t[r][c] = sum_q A_t[q][r] * B_t[q][c]
 * with 1-D factors  F_t[q][n] = sum_d W_t[q][d] / (z_n - e1[q][d] + i g1[q][d])
 *                                          [ / (z_n - e2[q][d] + i g2[q][d]) ]
 * ------------------------------------------------------------------------------------ */
/* d_z[n] real grid; d_W [T][R][D] complex; d_p1 [R][D][2] = (e1,g1); d_p2 same or NULL;
 * d_F [T][R][n] complex out */
int limeb200_sos_factor(const double* d_z, int n, const double* d_W, const double* d_p1,
                        const double* d_p2, int T, int R, int D, double* d_F, void* stream);
/* time-domain factors for the response functions of lime/signal/2DES.py:37-247:
 * F_t[q][n] = sum_d W_t[q][d] * (-i) theta(t_n) exp(-i e1[q][d] t_n - g1[q][d] t_n), theta(0) = 1 */
int limeb200_sos_factor_time(const double* d_t, int n, const double* d_W, const double* d_p1,
                             int T, int R, int D, double* d_F, void* stream);
/* d_out[T][nrow][ncol] (+)= scale * sum_q A[ta][q][row] * B[tb][q][col];
 * TA, TB = 1 (factor shared by all t) or T.                                               */
int limeb200_sos_outer(const double* d_A, int TA, const double* d_B, int TB, int T, int R,
                       int nrow, int ncol, double scale, int accumulate, double* d_out, void* stream);
/* TPA2D / TPA2D_time_order, lime/signal/sos.py:230-283: d_out[np][n1] real.
 * d_E[N], d_dip[N][N] real, d_gamma[N]; idx lists on device; time_order = 0/1             */
int limeb200_sos_tpa2d(const double* d_E, const double* d_dip, const double* d_gamma, int N,
                       const int* d_eidx, int ne, const int* d_fidx, int nf,
                       const double* d_omegap, int np, const double* d_omega1, int n1,
                       int time_order, double* d_out, void* stream);
/* ETPA double time integrals, _etpa lime/signal/sos.py:1171-1223: the second contraction
 *   d_out[c][f] = sum_a exp(i (d_Ef[f] - d_alpha[c]) d_t2[a]) d_V[a][c]      (complex d_V [n2][C], d_out [C][nf])
 * with V = (theta o (J + J^T)) . U1 from limeb200_zgemm, c = (pump frequency, intermediate state)              */
int limeb200_etpa_reduce(const double* d_V, const double* d_t2, int n2, const double* d_alpha, int C,
                         const double* d_Ef, int nf, double* d_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif
