/* lesgo_gpu.h -- C ABI of liblesgo_cuda.so: the B200 (sm_100a) implementation of
 * LESGO's per-timestep pseudo-spectral core, callable from Fortran through
 * ISO_C_BINDING (see fortran/ and INTEGRATION.md).
 *
 * Array layout everywhere: the Fortran layout of the reference, FP64, column major,
 * f(ld, ny, 0:nz) with ld = nx + 2, i.e. element (jx, jy, jz) (1-based x, y; z from
 * 0) at  f[(jz*ny + (jy-1))*ld + (jx-1)].  "nz" is the per-rank nz of param.f90 (each
 * array has nz+1 planes; the library is built for the MPI layout lbz = 0 only,
 * SURVEY Appendix A.2).  Arrays the reference declares 1:nz (dpdx, dpdy, dpdz) are
 * passed as the address of their plane 1 MINUS one plane, or simply allocate them
 * 0:nz -- every entry point documents which planes it reads and writes.
 * Complex values are interleaved along x (re, im).
 *
 * Pointers may be host pointers (the Fortran module arrays) or CUDA device pointers;
 * the library detects which (cudaPointerGetAttributes).  Host arrays are staged
 * through device buffers inside the call; device arrays are used in place.
 *
 * All functions return 0 on success, non-zero on error (lesgo_gpu_last_error gives
 * the text); the Fortran shim turns non-zero into `call error(...)`
 * (messages.f90:228-240).  There is NO CPU fallback: lesgo_gpu_create fails when no
 * CUDA device is usable.
 */
#ifndef LESGO_GPU_H
#define LESGO_GPU_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lesgo_gpu_ctx lesgo_gpu_ctx;

typedef struct lesgo_gpu_dims {
    int nx, ny, nz;          /* param.f90:96-99; nz = per-rank nz (levels 0..nz)        */
    int nz_tot;              /* (nz-1)*nproc + 1, input_util.f90:200                    */
    int nproc, coord;        /* z-slab decomposition, mpi_defs.f90:77-87                */
    double L_x, L_y, dz;     /* fft.f90:154-155 wavenumber scaling; input_util.f90:235  */
    int lbc_mom, ubc_mom;    /* wall types: select convec.f90:101-151 special planes    */
    int sgs;                 /* convec.f90:47-53: jzLo = 2 when LES, 1 when DNS         */
    int device;              /* CUDA device ordinal; -1 = current device; -2 - r = node-local
                                rank r, mapped to ordinal r mod (device count)          */
} lesgo_gpu_dims;

/* ---- lifetime (replaces init_fft, fft.f90:102-127) -------------------------------- */
int lesgo_gpu_create(const lesgo_gpu_dims* dims, lesgo_gpu_ctx** ctx);
int lesgo_gpu_destroy(lesgo_gpu_ctx* ctx);
const char* lesgo_gpu_last_error(const lesgo_gpu_ctx* ctx);   /* ctx may be NULL */
int lesgo_gpu_set_stream(lesgo_gpu_ctx* ctx, void* cuda_stream);
/* Page-lock (cudaHostRegister) / release a host array that is passed to the per-routine entry points again and
 * again -- the Fortran shim does it for the module arrays of sim_param (sim_param.f90:52-82), which are pageable
 * allocations: staged copies then run at PCIe rate and overlap with the kernels.  Idempotent. */
int lesgo_gpu_host_register(lesgo_gpu_ctx* ctx, void* host, size_t bytes);
int lesgo_gpu_host_unregister(lesgo_gpu_ctx* ctx, void* host);
int lesgo_gpu_synchronize(lesgo_gpu_ctx* ctx);
/* kernels launched by this context since creation (bench.py's gpu_launches) */
long lesgo_gpu_launch_count(const lesgo_gpu_ctx* ctx);
/* per-launch CUDA-event timing for bench.py / profiling: enable = 1/0; when report != NULL
 * it receives "label count total_ms" lines for the launches recorded so far (and clears them) */
int lesgo_gpu_profile(lesgo_gpu_ctx* ctx, int enable, char* report, int report_len);

/* ---- module fft (fft.f90) ------------------------------------------------------------ */
/* kx, ky, k2 (lh, ny) as init_wavenumber builds them, fft.f90:130-160 (host arrays) */
int lesgo_gpu_wavenumbers(lesgo_gpu_ctx* ctx, double* kx, double* ky, double* k2);
/* padd(u_big, u), fft.f90:43-71: nplanes planes of (ld, ny) -> (ld_big, ny2) */
int lesgo_gpu_padd(lesgo_gpu_ctx* ctx, double* u_big, const double* u, int nplanes);
/* unpadd(cc, cc_big), fft.f90:74-99 */
int lesgo_gpu_unpadd(lesgo_gpu_ctx* ctx, double* cc, const double* cc_big, int nplanes);
/* dfftw_execute_dft_r2c / c2r with the plans forw, back (big = 0) or forw_big,
 * back_big (big = 1), fft.f90:114-121; unnormalised, nplanes planes, out may equal in */
int lesgo_gpu_fft_r2c(lesgo_gpu_ctx* ctx, const double* in, double* out, int nplanes, int big);
int lesgo_gpu_fft_c2r(lesgo_gpu_ctx* ctx, const double* in, double* out, int nplanes, int big);

/* ---- FFTW3 legacy-Fortran API (the lower boundary) --------------------------------------
 * The library itself exports the symbols LESGO's remaining CPU routines call with the plan handles of
 * module fft or with plans of their own (fft.f90:114-121; test_filtermodule.f90:138-167; scalars.f90:513-624;
 * turbine_indicator.f90:130-151), so a LESGO build links WITHOUT libfftw3:
 *     dfftw_plan_dft_r2c_2d_(plan, n_fast, n_slow, in, out, flags)      dfftw_plan_dft_c2r_2d_(...)
 *     dfftw_execute_dft_r2c_(plan, in, out)   dfftw_execute_dft_c2r_(plan, in, out)   dfftw_destroy_plan_(plan)
 * all arguments by reference, plan = integer*8.  In-place plans of the bound context's (nx, ny) and
 * (3nx/2, 3ny/2) shapes run on the hot path's kernels; any other 2-3-5-smooth shape (e.g. the 2048 x 2048
 * out-of-place transforms of the disk indicator) runs on a generic device transform; anything else, or any
 * failure, prints and exits like `call error` (messages.f90:228-240).  The functions below are the same
 * operations with status returns (what the tests call): in-place real rows hold 2*(n_fast/2+1) doubles,
 * out-of-place real rows n_fast, complex rows n_fast/2+1 values (FFTW's r2c/c2r layout).
 * lesgo_gpu_create binds the context it creates (the last one wins); dims may be NULL when ctx is NULL. */
int lesgo_gpu_fftw_bind(lesgo_gpu_ctx* ctx, const lesgo_gpu_dims* dims);
int lesgo_gpu_fftw_plan_2d(int c2r, int n_fast, int n_slow, int inplace, long long* plan);
int lesgo_gpu_fftw_execute(long long plan, int c2r, double* in, double* out);
int lesgo_gpu_fftw_destroy(long long plan);
const char* lesgo_gpu_fftw_last_error(void);
void dfftw_plan_dft_r2c_2d_(long long* plan, const int* n_fast, const int* n_slow, double* in, double* out, const int* flags);
void dfftw_plan_dft_c2r_2d_(long long* plan, const int* n_fast, const int* n_slow, double* in, double* out, const int* flags);
void dfftw_execute_dft_r2c_(const long long* plan, double* in, double* out);
void dfftw_execute_dft_c2r_(const long long* plan, double* in, double* out);
void dfftw_destroy_plan_(long long* plan);

/* ---- module derivatives (derivatives.f90); all arrays (ld, ny, 0:nz) ------------------- */
int lesgo_gpu_ddx(lesgo_gpu_ctx* ctx, const double* f, double* dfdx);                   /* :37  */
int lesgo_gpu_ddy(lesgo_gpu_ctx* ctx, const double* f, double* dfdy);                   /* :79  */
int lesgo_gpu_ddxy(lesgo_gpu_ctx* ctx, const double* f, double* dfdx, double* dfdy);    /* :121 */
int lesgo_gpu_filt_da(lesgo_gpu_ctx* ctx, double* f, double* dfdx, double* dfdy);       /* :166, f inout */
int lesgo_gpu_ddz_uv(lesgo_gpu_ctx* ctx, const double* f, double* dfdz);                /* :214 */
int lesgo_gpu_ddz_w(lesgo_gpu_ctx* ctx, const double* f, double* dfdz);                 /* :267 */
/* test_filter(f) with a caller-supplied kernel G(lh, ny) (test_filtermodule.f90:126-146);
 * nplanes planes of (ld, ny), in place */
int lesgo_gpu_test_filter(lesgo_gpu_ctx* ctx, double* f, const double* G, int nplanes);

/* ---- convec (convec.f90:21-334); module arrays passed explicitly ----------------------- */
int lesgo_gpu_convec(lesgo_gpu_ctx* ctx, const double* u, const double* v, const double* w,
                     const double* dudy, const double* dudz, const double* dvdx,
                     const double* dvdz, const double* dwdx, const double* dwdy,
                     double* RHSx, double* RHSy, double* RHSz);

/* ---- press_stag_array (press_stag_array.f90:21-290) ------------------------------------
 * reads u, v, w (1:nz-1, + w(nz) on the top rank), divtz(1) on coord 0, divtz(nz) on the
 * top rank; writes p (0:nz), dpdx, dpdy, dpdz (all ADDRESSED as (ld, ny, 0:nz) arrays -- pass the
 * address of plane 1 minus one plane for the three 1:nz arrays; their plane 0 is never read, written
 * or copied; planes
 * 1:nz-1 valid, + nz of p/dpdz on the top rank).  Includes tridag_array and, for
 * nproc > 1, the halo / transpose communication (lesgo_gpu_comm_init first). */
int lesgo_gpu_press_stag_array(lesgo_gpu_ctx* ctx, const double* u, const double* v,
                               const double* w, const double* divtz, double dt, double tadv1,
                               double* p, double* dpdx, double* dpdy, double* dpdz);

/* ---- tridag_array (tridag_array.f90:166-246, serial form): general coefficients ---------
 * a, b, c (lh, ny, n), r, u (ld, ny, n), n rows; solves every (jx <= lh-1, jy /= ny/2+1,
 * (jx,jy) /= (1,1)) mode like the reference. */
int lesgo_gpu_tridag_array(lesgo_gpu_ctx* ctx, const double* a, const double* b,
                           const double* c, const double* r, double* u, int n);

/* ---- device-resident state (sim_param.f90:31-82) and whole-step entry --------------------- */
enum lesgo_gpu_field {
    LG_U = 0, LG_V, LG_W, LG_DUDX, LG_DUDY, LG_DUDZ, LG_DVDX, LG_DVDY, LG_DVDZ,
    LG_DWDX, LG_DWDY, LG_DWDZ, LG_RHSX, LG_RHSY, LG_RHSZ, LG_RHSX_F, LG_RHSY_F, LG_RHSZ_F,
    LG_P, LG_DPDX, LG_DPDY, LG_DPDZ, LG_DIVTX, LG_DIVTY, LG_DIVTZ,
    LG_TXX, LG_TXY, LG_TXZ, LG_TYY, LG_TYZ, LG_TZZ,
    /* Lagrangian scale-dependent model state (sgs_param.f90; lagrange_Sdep.f90, interpolag_Sdep.f90):
     * allocated only when a step with sgs_model = 5 (or an upload/download of them) asks for them */
    LG_F_LM, LG_F_MM, LG_F_QN, LG_F_NN, LG_CS_OPT2,
    /* applied body force of the actuator disks (sim_param.f90 fxa, fya, fza; forcing.f90:102-106):
     * allocated by lesgo_gpu_turbines_init */
    LG_FXA, LG_FYA, LG_FZA, LG_NFIELDS
};
/* device pointer of a resident field ((ld, ny, 0:nz) doubles); allocated on first use.  NOTE:
 * lesgo_gpu_step makes RHS* and RHS*_f trade places every step instead of copying (main.f90:155-157),
 * so re-query the pointers of those six fields after a step; upload/download always follow the ids. */
double* lesgo_gpu_field_ptr(lesgo_gpu_ctx* ctx, int field);
int lesgo_gpu_upload(lesgo_gpu_ctx* ctx, int field, const double* host);
int lesgo_gpu_download(lesgo_gpu_ctx* ctx, int field, double* host);

typedef struct lesgo_gpu_step_params {
    double dt, tadv1, tadv2;            /* main.f90:135-144, input_util.f90:403-408 */
    double mean_p_force_x, mean_p_force_y;   /* 0 when use_mean_p_force = .false.   */
    double ubot, utop, nu_molec_nd;     /* wallstress.f90 DNS walls: nu_molec/(z_i u_star) */
    int first_step;                     /* main.f90:273-280 Euler start                */
    int mode;                           /* 0 = core (divt* taken from the resident fields as given);
                                           1 = full step: + wallstress (lbc/ubc 0, 1, 2), calc_Sij,
                                           sgs_stag with a constant coefficient, divstress_uv/w   */
    /* mode 1 only (sgs_param.f90, sgs_stag_util.f90:87-189, wallstress.f90, test_filtermodule.f90) */
    int sgs_model;                      /* 1 = Smagorinsky + Mason wall damping; 5 = Lagrangian scale-dependent
                                           dynamic model (Cs_opt2 is the resident field LG_CS_OPT2, driven by the
                                           lasd_* members below); other: Cs_opt2 = 0.03, l = delta (the dynamic
                                           models before DYN_init)                                      */
    int ifilter;                        /* test filter of the equilibrium wall model: 1 cutoff, 2 Gaussian, 3 box */
    double Co, wall_damp_exp, vonk, zo; /* lesgo.conf MODEL / FLOW_COND                                  */
    /* sgs_model = 5 only.  The host keeps the step counters (jt, jt_total, DYN_init, cs_count, inilag) and
     * tells the step which branch of sgs_stag_util.f90:183-216 applies:
     *   lasd_cs_init: jt == 1 and inilag           -> Cs_opt2 = 0.03 everywhere (:187-189)
     *   lasd_update : jt >= DYN_init and mod(jt_total, cs_count) == 0 -> lagrange_Sdep() (:192-215), i.e.
     *                 interpolag_Sdep (semi-Lagrangian transport of F_LM, F_MM, F_QN, F_NN) + the 42 test
     *                 filters per plane + the running averages + Cs_opt2 (lagrange_Sdep.f90:22-430)
     *   lasd_init_F : inilag and (jt == cs_count or jt == DYN_init), first time: F_* initialised from
     *                 MM, NN (lagrange_Sdep.f90:270-281,320-331)
     *   lagran_dt   : sgs_stag_util.f90:73-82 (cs_count * dt for a fixed time step) */
    int lasd_cs_init, lasd_update, lasd_init_F;
    double lagran_dt;
    /* actuator disks (after lesgo_gpu_turbines_init): 1 = forcing_applied + main.f90:264-266 inside the step,
     * i.e. turbines_forcing on the velocities of time level n and RHS += (fxa, fya, fza); turbines_eps is the
     * time-filter weight (dt_dim / T_avg_dim) / (1 + dt_dim / T_avg_dim) of turbines.f90:563-567 */
    int turbines;
    double turbines_eps;
} lesgo_gpu_step_params;
/* One timestep main.f90:155-344 on the resident fields, no host round trip. */
int lesgo_gpu_step(lesgo_gpu_ctx* ctx, const lesgo_gpu_step_params* sp);
/* cfl_util.f90:35-69 get_max_cfl (dx, dy from L/n) and rmsdiv.f90 on resident fields */
int lesgo_gpu_max_cfl(lesgo_gpu_ctx* ctx, double dt, double* cfl);
/* get_cfl_dt, cfl_util.f90:72-113: the time step that makes the maximum CFL number equal `cfl` (min over ranks) */
int lesgo_gpu_cfl_dt(lesgo_gpu_ctx* ctx, double cfl, double* dt);
int lesgo_gpu_rmsdiv(lesgo_gpu_ctx* ctx, double* rms);

/* ---- restart file (io.f90:1173-1211 checkpoint, initial.f90:226-239 ic_file) ----------------------------
 * `fname` is this rank's file (the reference's checkpoint_file // '.c' // coord, e.g. "vel.out.c0"): one
 * Fortran sequential unformatted record with planes 1:nz of u, v, w, RHSx, RHSy, RHSz, Cs_opt2, F_LM, F_MM,
 * F_QN, F_NN (the model fields are zero when no dynamic model ran), native byte order, gfortran subrecords
 * above 2 GiB.  Written from / read into the resident fields; ghost planes are not part of the file, so
 * call lesgo_gpu_sync_real_array on u, v, w (device pointers) after a read when nproc > 1. */
int lesgo_gpu_checkpoint_write(lesgo_gpu_ctx* ctx, const char* fname);
int lesgo_gpu_checkpoint_read(lesgo_gpu_ctx* ctx, const char* fname);

/* ---- running time averages (time_average.f90:176-320, tavg%compute) -------------------------------------
 * Accumulates  acc += quantity * dt  for the 26 quantities of type tavg_t from the resident fields (u, v, w, p,
 * the stresses, the derivatives of the last step, Cs_opt2, the disk forces when turbines are initialised),
 * including the uv<->w grid interpolations and their ghost-plane exchanges.  tavg%finalize (division by
 * total_time, Reynolds stresses, file output, :323-480) stays with the host, which fetches an accumulator as
 * the (nx, ny, 0:nz) array tavg_t holds; total_time may be NULL. */
enum lesgo_gpu_tavg {
    LG_TA_U = 0, LG_TA_V, LG_TA_W, LG_TA_W_UV, LG_TA_U_W, LG_TA_V_W, LG_TA_U2, LG_TA_V2, LG_TA_W2, LG_TA_UV, LG_TA_UW,
    LG_TA_VW, LG_TA_TXX, LG_TA_TYY, LG_TA_TZZ, LG_TA_TXY, LG_TA_TXZ, LG_TA_TYZ, LG_TA_P, LG_TA_FX, LG_TA_FY, LG_TA_FZ,
    LG_TA_CS_OPT2, LG_TA_VORTX, LG_TA_VORTY, LG_TA_VORTZ, LG_TA_N
};
int lesgo_gpu_tavg_compute(lesgo_gpu_ctx* ctx, double dt);
int lesgo_gpu_tavg_download(lesgo_gpu_ctx* ctx, int which, double* host, double* total_time);
int lesgo_gpu_tavg_reset(lesgo_gpu_ctx* ctx);

/* ---- actuator-disk turbines (turbines.f90) ---------------------------------------------------------
 * The host keeps turbines_init / turbines_nodes (turbines.f90:129-462: input files, the filtered indicator
 * function of turbine_indicator.f90, node search) and hands the result over; call again when the disks
 * move (dyn_theta1/2).  use_rotation (turbines.f90:76, :607-615): lesgo_gpu_turbines_rotation below. */
typedef struct lesgo_gpu_turbine {
    int num_nodes;           /* wind_farm%turbine(s)%num_nodes on THIS rank (may be 0)                   */
    const int* nodes;        /* (num_nodes, 3) triplets i, j (1-based), k (local, 1..nz-1): %nodes(l,1:3)  */
    const double* ind;       /* (num_nodes) normalised indicator weights %ind(l)                          */
    double nhat[3];          /* unit normal, turbines.f90:344-346                                         */
    double Ct_prime, dia;
    double M;                /* %turb_ind_func%M, used when adm_correction                                */
    double u_d_T;            /* initial running-average disk velocity (turbine_vel_init, :642-676)        */
} lesgo_gpu_turbine;
int lesgo_gpu_turbines_init(lesgo_gpu_ctx* ctx, int nloc, const lesgo_gpu_turbine* turbines, int adm_correction);
/* ADM with rotation (use_rotation = .true., turbines.f90:607-615): after lesgo_gpu_turbines_init, hand over what
 * turbines_nodes (turbines.f90:419-429, :456) also leaves in wind_farm%turbine(s): ind_t[s] = %ind_t(1:num_nodes),
 * e_theta[s] = %e_theta(l, 1:3) as (num_nodes, 3) triplets like `nodes`; tip_speed_ratio as turbines.f90:78.  The
 * scatter then adds f_n * e_theta * ind_t / tip_speed_ratio to the three force components.  nloc must match the
 * init call; lesgo_gpu_turbines_init switches rotation off again. */
int lesgo_gpu_turbines_rotation(lesgo_gpu_ctx* ctx, int nloc, const double* const* ind_t, const double* const* e_theta,
                                double tip_speed_ratio);
/* turbines_forcing (turbines.f90:465-638) on the resident u, v, w -> resident LG_FXA, LG_FYA, LG_FZA (fza on
 * w nodes, ghost planes synchronised).  u_d, u_d_T, f_n: optional host arrays (nloc) receiving %u_d, %u_d_T,
 * %f_n after the update (NULL = leave everything on the device, no synchronisation). */
int lesgo_gpu_turbines_forcing(lesgo_gpu_ctx* ctx, double eps, double* u_d, double* u_d_T, double* f_n);

/* ---- multi-GPU (mpi_defs.f90 -> NCCL) --------------------------------------------------------
 * id: 128-byte ncclUniqueId made by lesgo_gpu_comm_unique_id on coord 0 and broadcast by
 * the host (MPI_Bcast in the Fortran shim, torch.distributed in the Python host). */
int lesgo_gpu_comm_unique_id(void* id128);
/* id of the SINGLE-DEVICE transport instead: every rank is a thread of this process and all contexts sit on the
 * same GPU (device-to-device copies ordered by events; NCCL refuses two ranks on one device).  It lets a one-GPU
 * box run -- and test -- the multi-slab path: halos, pressure transposes, k = 0 chain, reductions. */
int lesgo_gpu_comm_local_id(void* id128);
int lesgo_gpu_comm_init(lesgo_gpu_ctx* ctx, const void* id128);
/* Optional, one node: the two transposes of the pressure solve go over NVLink peer memory instead of NCCL
 * all-to-alls -- the right-hand-side assembly kernel stores straight into the pencil buffers of the other GPUs
 * and the unpacking kernel loads p_hat straight out of the buffers of the GPUs that solved it, so the transfers
 * cost no pass of their own.  Every rank exports a 128-byte blob, the host gathers them in rank order
 * (MPI_Allgather / torch.distributed) and every rank imports all nproc blobs (after lesgo_gpu_comm_init).
 * Ranks may be threads of one process (peer access) or separate processes (CUDA IPC).  Importing NULL switches
 * back to the NCCL all-to-alls (all ranks must agree: a host that sees the import fail on any rank disables it
 * on every rank). */
int lesgo_gpu_comm_p2p_export(lesgo_gpu_ctx* ctx, void* blob128);
int lesgo_gpu_comm_p2p_import(lesgo_gpu_ctx* ctx, const void* blobs128_times_nproc);
/* mpi_sync_real_array(var, 0, isync), mpi_defs.f90:167-264: isync 1 = DOWN, 2 = UP, 3 = DOWNUP */
int lesgo_gpu_sync_real_array(lesgo_gpu_ctx* ctx, double* var, int isync);

#ifdef __cplusplus
}
#endif
#endif /* LESGO_GPU_H */
