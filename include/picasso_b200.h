/*
 * picasso_b200.h -- C ABI of libpicasso_b200.so: the B200 (sm_100a) implementation
 * of picasso's single-molecule localization hot path.
 *
 * Conventions (the in-repo precedent for a native backend is Gpufit's ctypes
 * binding, reference picasso/ext/pygpufit/gpufit.py:37-76, 199-366):
 *   - every function returns an int status, 0 == PB_OK; on failure
 *     pb_last_error() returns a thread-local message (cf. gpufit_get_last_error);
 *   - the caller allocates every input and output as C-contiguous arrays and
 *     passes raw pointers + sizes; the library never keeps a pointer after the
 *     call returns;
 *   - plain `foo()` entry points take HOST pointers (pageable or pinned) and
 *     do the host<->device copies themselves, chunked and overlapped with the
 *     kernels; `foo_dev()` entry points take DEVICE pointers on the current
 *     CUDA device and are asynchronous on `stream` (a cudaStream_t passed as
 *     void*; NULL = default stream);
 *   - the library uses whichever device is current (cudaSetDevice / the
 *     caller's torch.cuda.set_device); one process per GPU.
 *
 * Each entry point names the reference interface it replaces (file:line in
 * jungmannlab/picasso @ 96e0da51).
 */
#ifndef PICASSO_B200_H
#define PICASSO_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PB_OK 0
#define PB_ERR_INVALID 1
#define PB_ERR_CUDA 2
#define PB_ERR_NOGPU 3
#define PB_ERR_CAPACITY 4
#define PB_ERR_CUFFT 5

/* status flag bits written per spot by pb_mle_fit* (status array, nullable) */
#define PB_MLE_FLAG_DEGENERATE_INIT 1 /* reference would raise ZeroDivisionError */
#define PB_MLE_FLAG_PINV_FALLBACK 2   /* Fisher matrix singular: eigen pseudo-inverse used */
#define PB_MLE_FLAG_NONFINITE 4       /* a CRLB entry is inf/nan */

/* ---- library / device management ------------------------------------- */
const char* pb_last_error(void);          /* cf. gpufit_get_last_error, gpufit.py:361-366 */
const char* pb_version(void);
int pb_device_count(void);                /* number of sm_100 devices visible (cf. gpufit cuda_available) */
int pb_set_device(int device);            /* cudaSetDevice for the calling thread */
int pb_get_device(int* device);           /* cudaGetDevice of the calling thread (worker threads inherit it
                                             through the Python layer: the CUDA current device is per thread) */
int pb_synchronize(void);                 /* cudaDeviceSynchronize on the current device */
/* pinned host memory for callers that want full PCIe speed (optional) */
int pb_host_alloc(void** ptr, size_t bytes);
int pb_host_free(void* ptr);
/* Transfers between ordinary (pageable) host memory and device buffers at close to PCIe speed:
 * chunks alternate between two pinned staging buffers filled / drained by PB_COPY_THREADS host
 * threads (default min(12, cores)); pinned host memory is copied directly.  pb_copy_h2d returns
 * once the last chunk is enqueued on `stream` (a pageable h_src may be reused immediately);
 * pb_copy_d2h is blocking.  Every host-pointer entry point below uses these internally. */
int pb_copy_h2d(void* d_dst, const void* h_src, size_t bytes, void* stream);
int pb_copy_d2h(void* h_dst, const void* d_src, size_t bytes, void* stream);
/* Peer-to-peer gather plumbing (one process per GPU on an NVLink box, csrc/p2p.cu): device
 * buffers allocated here can be exported as CUDA IPC handles (pb_ipc_handle_bytes() bytes),
 * opened by the other ranks and written with pb_copy_d2d_async -- copy-engine transfers that
 * take no SM from a running fit, unlike an NCCL all-gather kernel. */
int pb_dev_alloc(void** ptr, size_t bytes);
int pb_dev_free(void* ptr);
int pb_ipc_handle_bytes(void);
int pb_ipc_export(const void* d_ptr, void* handle);
int pb_ipc_open(const void* handle, void** d_ptr);
int pb_ipc_close(void* d_ptr);
int pb_copy_d2d_async(void* dst, const void* src, size_t bytes, void* stream);
/* number of kernel launches issued by this library in this process (bench.py's gpu_launches) */
long long pb_launch_count(void);

/* ---- MLE Gaussian fit --------------------------------------------------
 * Replaces picasso.gaussmle.gaussmle / gaussmle_async and the numba kernels
 * _mlefit_sigmaxy / _mlefit_sigma (+ _crlb)      picasso/gaussmle.py:409-954.
 *   spots      (n, box, box) float32, photons
 *   eps        convergence criterion (Python float -> double)
 *   max_it     maximum Newton iterations (0 => theta = start values)
 *   method     0 = "sigma", 1 = "sigmaxy"; anything else -> PB_ERR_INVALID
 *              with message "Method not available." (gaussmle.py:465)
 *   thetas     (n, 6) float32  [x, y, photons, bg, sx, sy]
 *   crlbs      (n, 6) float32
 *   logliks    (n,)   float32
 *   iterations (n,)   int32
 *   status     (n,)   int32 flag bits, may be NULL
 *   progress   nullable; the host variant stores the number of spots finished
 *              so far (the reference's `current[0]`, gaussmle.py:386-406)
 * box: odd, 5..21.
 */
int pb_mle_fit(size_t n, int box, const float* spots, double eps, int max_it, int method,
               float* thetas, float* crlbs, float* logliks, int* iterations, int* status,
               volatile long long* progress);
int pb_mle_fit_dev(size_t n, int box, const float* d_spots, double eps, int max_it, int method,
                   float* d_thetas, float* d_crlbs, float* d_logliks, int* d_iterations,
                   int* d_status, void* stream);

/* Kernel selection for pb_mle_fit* (tuning / A-B measurement; results agree within the parity
 * bar for all three): 0 = lane-group kernel (8/16/32 lanes per spot, csrc/mle_fit.cu),
 * 1 = thread-per-spot kernels with float64 edge terms and per-pixel sums, 2 = thread-per-spot with a
 * table-driven float64 PSF, float32 derivative factors and float32 per-pixel sums over a float-float
 * residual (csrc/mle_tps.cu; default for box <= 13).  PB_MLE_IMPL in the environment
 * sets the initial value.  Boxes above 13 always use the lane-group kernel. */
int pb_mle_set_impl(int impl);
int pb_mle_get_impl(void);

/* Measurement hook for bench.py: when enabled, the thread-per-spot path records CUDA events on
 * the launch stream around its three kernels; pb_mle_profile_read waits for the most recent
 * call and returns their durations in ms: {start values, Newton iterations, CRLB + logL}. */
int pb_mle_profile(int enable);
/* Phase hook for multi-GPU callers: a cudaEvent_t (NULL clears it) that the following thread-per-spot fits
 * record on their launch stream after the iteration kernel: thetas and iterations are final from there on
 * and can be sent to the peers while the CRLB kernel runs. */
int pb_mle_set_phase_event(void* cuda_event);
int pb_mle_profile_read(float* ms3);

/* ---- spot identification ------------------------------------------------
 * Replaces localize.identify_in_image / identify_in_frame / identify_by_frame_number
 * and the frame loops _identify_serial / identify_async
 * (picasso/localize.py:97-337, 340-421, 482-636).
 *   movie        (n_frames, Y, X) uint16 (dtype 0) or float32 (dtype 1)
 *   frame_offset number of movie[0] in the full movie (added to emitted frames)
 *   box          odd, 3..15;  min_ng: keep spots with net_gradient > min_ng
 *   roi          nullable int[4] = {y_start, x_start, y_end, x_end}: the frame is
 *                sliced first (python slice semantics), coordinates are offset back
 *   frame,x,y    int64 outputs, ng float32; capacity = their length
 *   n_found      number of spots found; if it exceeds capacity the call returns
 *                PB_ERR_CAPACITY and the caller retries with a larger buffer
 * The host variant returns the spots sorted by (frame, y, x) -- the order of the
 * serial reference; the _dev variant appends in arbitrary order and only
 * advances *d_counter (a device uint64 the caller zeroes).
 * Frame-bounds filtering (localize.py:395-409) is done by the caller by
 * choosing which frames to pass. */
#define PB_DTYPE_U16 0
#define PB_DTYPE_F32 1
int pb_identify(const void* movie, int dtype, size_t n_frames, int Y, int X,
                long long frame_offset, int box, double min_ng, const int* roi,
                long long* frame, long long* x, long long* y, float* ng, size_t capacity,
                size_t* n_found);
int pb_identify_dev(const void* d_movie, int dtype, size_t n_frames, int Y, int X,
                    long long frame_offset, int box, double min_ng, const int* roi,
                    long long* d_frame, long long* d_x, long long* d_y, float* d_ng,
                    size_t capacity, unsigned long long* d_counter, void* stream);

/* ---- ROI extraction + photon conversion ------------------------------------
 * Replaces localize.get_spots = _cut_spots(_numba/_framebyframe) + _to_photons
 * (picasso/localize.py:917-1145): spots[i] = (f32(movie[frame_i, y_i-r:y_i+r+1,
 * x_i-r:x_i+r+1]) - baseline) * sensitivity / gain, float32 arithmetic.
 * `movie` holds frames [frame_offset, frame_offset + n_frames); spots of other
 * frames are left untouched (callers stream a memmapped movie chunk by chunk). */
int pb_get_spots(const void* movie, int dtype, size_t n_frames, int Y, int X,
                 long long frame_offset, size_t n, const long long* frame, const long long* x,
                 const long long* y, int box, float baseline, float sensitivity, float gain,
                 float* spots);
int pb_get_spots_dev(const void* d_movie, int dtype, size_t n_frames, int Y, int X,
                     long long frame_offset, size_t n, const long long* d_frame,
                     const long long* d_x, const long long* d_y, int box, float baseline,
                     float sensitivity, float gain, float* d_spots, void* stream);

/* Fused identify + get_spots over a host movie chunk (one upload): the end-to-end
 * localize path, localize.py:1787-1811 (identify -> fit2D -> get_spots).  Outputs as
 * pb_identify (sorted by frame, y, x) plus spots (capacity, box, box) float32. */
int pb_identify_get_spots(const void* movie, int dtype, size_t n_frames, int Y, int X,
                          long long frame_offset, int box, double min_ng, const int* roi,
                          float baseline, float sensitivity, float gain, long long* frame,
                          long long* x, long long* y, float* ng, float* spots, size_t capacity,
                          size_t* n_found);

/* ---- fused movie -> localisation table -----------------------------------------
 * Replaces the body of localize.localize (picasso/localize.py:1682-1815): identify
 * (:639-712) -> get_spots (:1115-1145) -> gaussmle.gaussmle (gaussmle.py:409-475) or
 * gausslq.fit_spots (gausslq.py:247-289) -> locs_from_fits (gaussmle.py:957-1037,
 * gausslq.py:404-544) in ONE pass over a host movie chunk: every frame crosses PCIe once,
 * identifications / ROIs / theta / CRLB stay on the GPU, only the finished localisation
 * columns come back.
 *   fit      0 = MLE "sigma", 1 = MLE "sigmaxy", 2 = LQ, 3 = LQ in the Gpufit column layout
 *            (fit_spots_gpufit + locs_from_fits_gpufit, gausslq.py:346-395, 487-544)
 *   em       camera_info["Gain"] > 1 (doubles the LQ localisation variance, gausslq.py:547-589)
 *   columns  (pb_locs_columns(fit), capacity) 4-byte elements, column-major:
 *            MLE (17): frame u32, x, y, photons, sx, sy, bg, lpx, lpy, ellipticity, net_gradient,
 *                      log_likelihood f32, iterations u32, photons_unc, bg_unc, sx_unc, sy_unc f32
 *            LQ  (11): frame u32, x, y, photons, sx, sy, bg, lpx, lpy, ellipticity, net_gradient
 *            rows ordered by (frame, y, x) of the identification -- the serial reference's order
 *   n_found  number of localisations; above `capacity` the call returns PB_ERR_CAPACITY with
 *            the required capacity in *n_found
 * Pageable movies are staged through pinned buffers by PB_COPY_THREADS (default 8) host
 * threads; pinned movies (pb_host_alloc) are copied directly. */
int pb_locs_columns(int fit);
int pb_localize(const void* movie, int dtype, size_t n_frames, int Y, int X, long long frame_offset,
                int box, double min_ng, const int* roi, float baseline, float sensitivity,
                float gain, int fit, double eps, int max_it, int em, void* columns,
                size_t capacity, size_t* n_found);
/* Same with the movie and the column block resident in HBM (multi-GPU frame shards, benchmarks
 * without PCIe); synchronous: the identification count of every chunk is read back. */
int pb_localize_dev(const void* d_movie, int dtype, size_t n_frames, int Y, int X,
                    long long frame_offset, int box, double min_ng, const int* roi, float baseline,
                    float sensitivity, float gain, int fit, double eps, int max_it, int em,
                    void* d_columns, size_t capacity, size_t* n_found);
/* The column arithmetic alone (numpy's evaluation order and dtypes, float32 IEEE operations):
 * identifications + fit results -> the columns above.  crlbs / logliks / iterations are read
 * for fit 0/1 only; for fit 3 `thetas` is in the Gpufit layout [photons, x, y, sx, sy, bg].
 * `ld` = elements between two columns of d_columns (>= n). */
int pb_locs_from_fits(size_t n, int fit, int box, int em, const long long* frame,
                      const long long* x, const long long* y, const float* ng, const float* thetas,
                      const float* crlbs, const float* logliks, const int* iterations,
                      void* columns);
int pb_locs_from_fits_dev(size_t n, int fit, int box, int em, const long long* d_frame,
                          const long long* d_x, const long long* d_y, const float* d_ng,
                          const float* d_thetas, const float* d_crlbs, const float* d_logliks,
                          const int* d_iterations, void* d_columns, size_t ld, void* stream);

/* ---- least-squares Gaussian fit --------------------------------------------
 * Replaces picasso.gausslq.fit_spot / fit_spots / fit_spots_parallel
 * (picasso/gausslq.py:206-343: scipy.optimize.leastsq == MINPACK lmdif with
 * ftol = xtol = 1e-2 on a float32-rounded point-sampled Gaussian) and stands in
 * for the vendored Gpufit DLL call gpufit_constrained (ext/pygpufit/gpufit.py:40-61,
 * gausslq.py:346-395).
 *   spots  (n, box, box) float32;  box odd, 5..15
 *   thetas (n, 6) float32 [x, y, photons, bg, sx, sy], x/y relative to the box centre
 *   infos  (n,) int32 MINPACK info codes (1-4 converged), nullable
 *   nfevs  (n,) int32 residual evaluations used, nullable */
int pb_lq_fit(size_t n, int box, const float* spots, float* thetas, int* infos, int* nfevs);
int pb_lq_fit_dev(size_t n, int box, const float* d_spots, float* d_thetas, int* d_infos,
                  int* d_nfevs, void* stream);
/* measurement hook (process-wide): 0 = register-resident factorisation of J^T J (default),
 * 1 = MINPACK-order Householder qrfac on the m x 6 Jacobian (A/B parity and speed runs) */
int pb_lq_set_impl(int impl);

/* ---- Gpufit-path least-squares fit ("gausslq-gpu") -----------------------------------------
 * Replaces the call into the vendored Gpufit 1.2.0 DLL: picasso.gausslq.fit_spots_gpufit
 * (picasso/gausslq.py:346-395) = _initial_parameters_gpufit (:128-148) + gpufit_fit
 * (picasso/ext/pygpufit/gpufit.py:40-61; GAUSS_2D_ELLIPTIC, LSE, tolerance, max iterations) +
 * amplitude * 2 pi sx sy (:393).  Gpufit's published float32 LM algorithm, one thread per fit
 * (csrc/gpufit_core.cuh); parity with the Windows binary is unpinned.
 *   spots   (n, box, box) float32, box odd 5..15
 *   params  (n, 6) float32 [photons, x, y, sx, sy, bg], x / y in pixel indices of the box
 *   states  (n) int32 nullable: 0 converged, 1 max iterations, 2 singular Hessian (Gpufit's codes)
 *   chi2    (n) float32 nullable;  n_iterations (n) int32 nullable */
int pb_gpufit_fit(size_t n, int box, const float* spots, float tolerance, int max_iterations,
                  float* params, int* states, float* chi2, int* n_iterations);
int pb_gpufit_fit_dev(size_t n, int box, const float* d_spots, float tolerance, int max_iterations,
                      float* d_params, int* d_states, float* d_chi2, int* d_n_iterations, void* stream);

/* ---- astigmatic z fit ----------------------------------------------------------
 * Replaces the per-localization loop and column arithmetic of picasso.zfit._fit_z
 * (picasso/zfit.py:327-383, behind zfit.zfit :465-646 / localize.localize_3D
 * localize.py:1818-2034): scipy.optimize.minimize_scalar(_fit_z_target, bounds=[-1000, 1000])
 * = scipy's bounded Brent minimiser (xatol 1e-5, maxiter 500) on the target of zfit.py:255-291
 * with the degree-6 calibration polynomials cx, cy (7 float64 each, z^6 .. z^0), then
 *   z = float32(x_min) * magnification,  d_zcalib = sqrt(float32(f_min)),
 *   lpz = _axial_localization_precision_astig (zfit.py:805-890), float32 like the reference's
 *         pandas columns; sigma errors from gausslq.sigma_uncertainty (method 0),
 *         gaussmle.sigma_uncertainty (1) or the sx_unc / sy_unc columns (2).
 *   sx, sy, photons, bg, sx_unc, sy_unc  float32[n] (photons/bg only for lpz, *_unc only for
 *   method 2);  z, d_zcalib, lpz float32[n] (lpz nullable);  nfev int32[n] nullable.
 * ensure_sanity / filter_z_fits stay on the host (picasso_b200/zfit.py). */
int pb_zfit(size_t n, const float* sx, const float* sy, const float* photons, const float* bg,
            const float* sx_unc, const float* sy_unc, const double* cx, const double* cy,
            double magnification, double pixelsize, int method, float* z, float* d_zcalib,
            float* lpz, int* nfev);
int pb_zfit_dev(size_t n, const float* d_sx, const float* d_sy, const float* d_photons,
                const float* d_bg, const float* d_sx_unc, const float* d_sy_unc, const double* cx,
                const double* cy, double magnification, double pixelsize, int method, float* d_z,
                float* d_d_zcalib, float* d_lpz, int* d_nfev, void* stream);

/* ---- NVSwitch multicast (NVLS) buffers and the fused fit + all-gather ------------------------
 * the sharded fit exchanges one thing, the all-gather of every rank's packed output
 * block.  A multicast object binds memory of all N GPUs at the same offsets; a store to the
 * multicast mapping is replicated by the switch into every GPU's copy.  One process per GPU:
 *   rank 0        pb_mc_create -> POSIX file descriptor, passed to the other ranks (Unix socket)
 *   other ranks   pb_mc_import(fd)
 *   all ranks     pb_mc_add_device; BARRIER; pb_mc_bind_map -> uc_ptr (this GPU's copy, ordinary
 *                 pointer) and mc_ptr (multicast mapping, store-only); BARRIER before first use
 * pb_mle_fit_gather_dev is pb_mle_fit_dev whose finishing kernel also stores each spot's 14 output
 * words through `mc_block` = mc_ptr + this rank's block offset, block layout
 * [thetas 6n | crlbs 6n | logliks n | iterations n] (n even): compute and collective in one kernel.
 * pb_mc_copy_async pushes an existing device buffer through the mapping (a few CTAs). */
int pb_mc_supported(void);
int pb_mc_padded_size(size_t bytes, int n_devices, size_t* padded);
int pb_mc_create(size_t padded_bytes, int n_devices, void** handle, int* export_fd);
int pb_mc_import(int fd, size_t padded_bytes, int n_devices, void** handle);
int pb_mc_add_device(void* handle);
int pb_mc_bind_map(void* handle, void** uc_ptr, void** mc_ptr);
int pb_mc_destroy(void* handle);
int pb_mc_copy_async(void* mc_dst, const void* d_src, size_t bytes, int n_ctas, void* stream);
int pb_mle_fit_gather_dev(size_t n, int box, const float* d_spots, double eps, int max_it, int method,
                          float* d_thetas, float* d_crlbs, float* d_logliks, int* d_iterations,
                          int* d_status, void* mc_block, void* stream);
/* which kernel emits what (process-wide): 1 = all 14 words from the finishing (CRLB) kernel; 2 = thetas +
 * iterations from the iteration kernel as lanes finish -- spread over the whole step, which keeps the
 * inbound NVLink rate of an N-rank gather low -- while the caller pushes the CRLB / logL half of its
 * block ([6n, 13n) floats) through the mapping behind the next step */
int pb_mle_fit_gather_mode(int mode);

/* ---- localisation table: ensure_sanity, z-fit filter, record packing on the device ----------
 * Replaces lib.ensure_sanity (picasso/lib.py:1786-1832), the tail of zfit._fit_z (picasso/zfit.py:
 * 356-383: append z / d_zcalib / lpz, ensure_sanity, filter_z_fits :675-704) and the record packing
 * of io.save_locs (picasso/io.py:2089-2110, locs.to_records(index=False)) for tables whose columns
 * are all 4 bytes wide (float32 / uint32 / int32 -- every table the fit path produces).
 *   cols[k]      host pointer to column k (n elements); is_float[k] != 0: float32 (rows with inf /
 *                NaN are dropped)
 *   ix, iy       indices of the x / y columns (x < width, y < height in float32), -1 if absent
 *   nonneg_cols  columns that must be >= 0 (x, y, lpx, lpy, lpz, photons, ellipticity, sx, sy
 *                as present)
 *   zfit         nullable: run the astigmatic z fit (pb_zfit_dev) on the resident columns first;
 *                z, d_zcalib, lpz become columns ncols .. ncols + 2 (lpz joins the >= 0 list) and,
 *                for filter_range > 0, rows with d_zcalib > range * sqrt(nanmean(d_zcalib^2)) are
 *                dropped after the sanity pass -- the RMSD reproduces numpy's float32 pairwise
 *                summation bit for bit
 *   out          out_records == 0: (ncols [+3], capacity) column block; != 0: n_kept packed records
 *                of ncols [+3] 4-byte fields (the layout of DataFrame.to_records(index=False))
 *   kept_index   nullable: original row numbers of the kept rows (capacity int64)
 *   n_kept       rows kept; above `capacity` the call returns PB_ERR_CAPACITY with the number */
typedef struct PbZfitSpec {
    int i_sx, i_sy, i_photons, i_bg, i_sx_unc, i_sy_unc;   /* column indices (-1: absent) */
    double cx[7], cy[7];                                    /* calibration polynomials z^6 .. z^0 */
    double magnification, pixelsize;
    int method;                                             /* as pb_zfit */
    int filter_range;                                       /* filter_z_fits range; 0 = no filter */
} PbZfitSpec;
int pb_locs_filter(size_t n, int ncols, const void* const* cols, const int* is_float, int ix, int iy,
                   const int* nonneg_cols, int n_nonneg, double width, double height,
                   const PbZfitSpec* zfit, int out_records, void* out, size_t capacity,
                   long long* kept_index, size_t* n_kept);

/* ---- AIM drift correction: intersection counting ---------------------------------
 * Replaces the counting core of picasso.aim (picasso/aim.py): _point_intersect_2d :297-344,
 * _point_intersect_3d :377-431, _run_intersections(_multithread) :148-266 and
 * _count_intersections :89-126 inside intersection_max :517-659 / intersection_max_z :662-773.
 * A handle keeps the target coordinates, the reference hash table (1-D int32 index ->
 * multiplicity) and scratch tables on the GPU for one round:
 *   pb_aim_set_targets    localizations in frame order; each array float32 (flag 0) or float64
 *                         (flag 1) -- the index arithmetic is done in the array's own dtype
 *                         like the reference's pandas columns; z nullable (2-D)
 *   pb_aim_set_reference  quantises the reference coordinates (index = round(x/d) +
 *                         round(y/d)*width_units [+ round(z/d)*width_units*height_units],
 *                         np.int32 truncation) and builds the table; rz == NULL selects 2-D
 *   pb_aim_count          targets [first, first+count): adds rel_x/rel_y (2-D) or rel_z (3-D)
 *                         before quantisation, then for each of the n_shifts shifts
 *                         roi_cc[j] = sum_c min(ref[c + shift_j], target[c]).  2-D shifts are
 *                         int32 values (int32 wrap-around sum), 3-D shifts float64.
 * The sub-pixel peak, spline and subtraction stay on the host (picasso_b200/aim.py). */
int pb_aim_create(void** handle);
int pb_aim_destroy(void* handle);
int pb_aim_set_targets(void* handle, size_t n, const void* x, int x_f64, const void* y, int y_f64,
                       const void* z, int z_f64);
int pb_aim_set_reference(void* handle, size_t n_ref, const void* rx, int rx_f64, const void* ry,
                         int ry_f64, const void* rz, int rz_f64, double intersect_d,
                         double width_units, double height_units);
int pb_aim_count(void* handle, size_t first, size_t count, double rel_x, double rel_y, double rel_z,
                 int n_shifts, const double* shifts, int* roi_cc);

/* ---- linking localizations into binding events -------------------------------------
 * Replaces the numba loops behind picasso.postprocess.link (picasso/postprocess.py:2007-2072):
 * pb_link_groups = _get_link_groups :2440-2507 (+ _get_next_loc_index_in_link_group :2510-2552):
 *   frame (sorted ascending, int64), x / y (float32, or float64 when xy_f64), group int32,
 *   d_max, max_dark_time -> link_group int32[n] (0 .. n_groups-1 in the reference's order of
 *   chain starts) and *n_groups.  Bit-exact with the sequential reference.
 * pb_link_reduce = _link_group_sum / _link_group_min_max / _link_group_last :2567-2661 for
 *   several columns at once: per group, in localization order, in the column's dtype
 *   (dtype 0 f32, 1 f64, 2 u32, 3 i32; op 0 sum, 1 min, 2 max, 3 last); outs[k] has n_groups
 *   elements of the column's dtype.  link_group must use every value 0 .. n_groups-1. */
int pb_link_groups(size_t n, const long long* frame, const void* x, const void* y, int xy_f64,
                   const int* group, double d_max, long long max_dark_time, int* link_group,
                   int* n_groups);
int pb_link_reduce(size_t n, const int* link_group, int n_groups, int n_cols,
                   const void* const* cols, const int* dtypes, const int* ops, void* const* outs);

/* ---- rendering ----------------------------------------------------------------
 * Replaces the unrotated paths of picasso.render.render (picasso/render.py:37-174):
 * _render_hist (:798-853, mode 0), _render_gaussian (:1020-1112, mode 1) and
 * _render_gaussian_iso (:1148-1216, mode 2) with _render_setup (:177-232),
 * _fill (:451-467) and _draw_gaussian_loc/_fill_gaussian (:494-575).
 *   x, y, lpx, lpy  float32[n] localisation coordinates / precisions (camera px);
 *                   lpx/lpy are ignored (may be NULL) for mode 0
 *   oversampling    display pixels per camera pixel
 *   viewport        (y_min, x_min) .. (y_max, x_max), strict inequalities
 *   image           float32 [n_pixel_y][n_pixel_x], zeroed by the call;
 *                   n_pixel = ceil(oversampling * (max - min)) is computed by the caller
 *   n_in_view       number of localisations inside the viewport (the reference's `n`)
 * The _dev variant takes an optional workspace (pb_render_workspace_bytes) enabling
 * the tile-binned shared-memory path; without it the direct-atomics path is used.
 * Unknown mode -> PB_ERR_INVALID, message "blur_method not understood." (render.py:174). */
size_t pb_render_workspace_bytes(size_t n, int n_pixel_y, int n_pixel_x);
/* Accumulation pass of the binned path (measurement hook, like pb_mle_set_impl): 0 = 64x64 tiles with
 * shared-memory atomics (default), 1 = warp-owned strips without shared-memory atomics (measured
 * slower; csrc/render.cu).  Same results within float32 summation order. */
int pb_render_set_impl(int impl);
int pb_render_get_impl(void);
int pb_render(size_t n, const float* x, const float* y, const float* lpx, const float* lpy,
              double oversampling, double y_min, double x_min, double y_max, double x_max,
              double min_blur_width, int mode, float* image, int n_pixel_y, int n_pixel_x,
              long long* n_in_view);
int pb_render_dev(size_t n, const float* d_x, const float* d_y, const float* d_lpx,
                  const float* d_lpy, double oversampling, double y_min, double x_min, double y_max,
                  double x_max, double min_blur_width, int mode, float* d_image, int n_pixel_y,
                  int n_pixel_x, unsigned long long* d_count, void* d_workspace,
                  size_t workspace_bytes, void* stream);
/* Multi-GPU rendering by image row bands (SURVEY.md 8e option B; render.py:1020 _render_gaussian is
 * what is being sharded).  pb_render_band_dev renders rows [row0, row0 + n_rows) of the image:
 * d_image holds n_rows x n_pixel_x floats, windows are clipped to the band, *d_count counts the
 * in-view localisations whose centre row lies in the band (the counts of disjoint bands add up to
 * the reference's n; the bands concatenate to the full image).  pb_render_band_count_dev /
 * pb_render_band_scatter_dev bucket a rank's share of the localisations by destination band
 * (band b = rows [band_rows[b], band_rows[b+1]), host array of n_bands + 1 ints, <= 64 bands): a
 * localisation goes to every band its 3-sigma window can reach; the scatter writes band b's
 * (x, y, lpx, lpy) float4 records from record index d_offsets[b] of the send buffer (the exchange
 * itself is ONE NCCL all-to-all of the records). */
int pb_render_band_dev(size_t n, const float* d_x, const float* d_y, const float* d_lpx,
                       const float* d_lpy, double oversampling, double y_min, double x_min,
                       double y_max, double x_max, double min_blur_width, int mode, float* d_image,
                       int n_pixel_y, int n_pixel_x, int row0, int n_rows,
                       unsigned long long* d_count, void* d_workspace, size_t workspace_bytes,
                       void* stream);
int pb_render_band_count_dev(size_t n, const float* d_x, const float* d_y, const float* d_lpx,
                             const float* d_lpy, double oversampling, double y_min, double x_min,
                             double y_max, double x_max, double min_blur_width, int mode,
                             int n_pixel_y, int n_pixel_x, int n_bands, const int* band_rows,
                             unsigned long long* d_counts, void* stream);
int pb_render_band_scatter_dev(size_t n, const float* d_x, const float* d_y, const float* d_lpx,
                               const float* d_lpy, double oversampling, double y_min, double x_min,
                               double y_max, double x_max, double min_blur_width, int mode,
                               int n_pixel_y, int n_pixel_x, int n_bands, const int* band_rows,
                               const unsigned long long* d_offsets, unsigned long long* d_cursor,
                               float* d_records, void* stream);
/* (x, y, lpx, lpy) float4 records (what the scatter writes and the all-to-all moves) -> columns */
int pb_render_unpack_records_dev(size_t n, const float* d_records, float* d_x, float* d_y,
                                 float* d_lpx, float* d_lpy, void* stream);

/* ---- RCC cross-correlation ----------------------------------------------------
 * Replaces the FFT work of picasso.imageprocess.xcorr / get_image_shift / rcc
 * (picasso/imageprocess.py:27-217) inside postprocess.undrift
 * (picasso/postprocess.py:2903-2961): for every pair i<j,
 *   window = crop(fftshift(real(ifft2(fft2(seg_i) * conj(fft2(seg_j))))) / sqrt(Y*X))
 * with the crop rows [Y0, Y0+H) and columns [X0, X0+W) (the reference's centre crop,
 * imageprocess.py:88-101; Y0 = X0 = 0, H = Y, W = X gives the whole correlation).
 *   segments (n_seg, Y, X) float32;  windows (n_seg*(n_seg-1)/2, H, W) float32 in the
 *   reference's pair order (i outer, j inner);  sums (n_seg) float64 = np.sum(seg)
 *   (the reference returns (0,0) for pairs with a zero-sum image, :83-84).
 * The arg-max / 5x5 peak fit / minimize_shifts stay on the host. */
int pb_rcc_windows(int n_seg, int Y, int X, const float* segments, int Y0, int X0, int H, int W,
                   float* windows, double* sums);
/* Inverse-transform strategy of the pair stage: -1 = auto (default: pruned inverse DFT of the
 * window rows / columns only when the window covers at most a quarter of the image, else cuFFT
 * C2R + crop), 0 = always cuFFT, 1 = always pruned.  PB_RCC_PRUNED=0/1 in the environment sets
 * the initial value. */
int pb_rcc_set_mode(int mode);
/* building blocks with device pointers (multi-GPU callers shard pairs across ranks) */
int pb_rcc_spectra_dev(int n_seg, int Y, int X, const float* d_segments, void* d_spectra,
                       double* d_sums, void* stream);
int pb_rcc_windows_dev(int n_pairs, const int* d_pair_i, const int* d_pair_j, int Y, int X,
                       const void* d_spectra, int Y0, int X0, int H, int W, float* d_windows,
                       int batch, void* d_workspace, size_t workspace_bytes, void* stream);
/* the per-pair body of get_image_shift (imageprocess.py:103-157) on device windows: 32 float64 per
 * pair = {status (0 fitted, 1 refit on the host, 2 window touches the crop edge -> (0,0), 3 odd
 * cut-out), arg-max y, arg-max x, xc, yc, the 5 x 5 window, padding} */
int pb_rcc_peakfit_dev(int n_pairs, const float* d_windows, int H, int W, double* d_records,
                       void* stream);
/* segments per L2 pair tile for images of Y rows (pairs sorted by (i / TS, j / TS) run tile by tile) */
int pb_rcc_tile_segments(int Y);

/* Fused front end of postprocess.undrift (picasso/postprocess.py:2903-2961 = segment
 * :2846-2900 + imageprocess.rcc :160-217): segment images are rendered on the device
 * (oversampling 1, full FOV, blur "gaussian", min_blur_width as given) from localisations
 * grouped by segment -- segment i owns [seg_start[i], seg_start[i+1]) of x/y/lpx/lpy --
 * and cross-correlated without leaving the GPU.  Outputs as pb_rcc_windows; segments_out
 * (n_seg, Y, X) float32 is optional. */
int pb_undrift_windows(int n_seg, const long long* seg_start, const float* x, const float* y,
                       const float* lpx, const float* lpy, int Y, int X, double min_blur_width,
                       int Y0, int X0, int H, int W, float* windows, double* sums,
                       float* segments_out);
/* Same for a subset of the pairs (pair_i[k], pair_j[k]), k < n_pairs; windows (n_pairs, H, W).
 * Multi-GPU undrift: every rank renders and transforms all segments (cheap) and correlates
 * only its share of the pairs -- no exchange of spectra.  n_pairs < 0 selects all i < j pairs. */
int pb_undrift_windows_pairs(int n_seg, const long long* seg_start, const float* x, const float* y,
                             const float* lpx, const float* lpy, int Y, int X, double min_blur_width,
                             int Y0, int X0, int H, int W, int n_pairs, const int* pair_i,
                             const int* pair_j, float* windows, double* sums, float* segments_out);

/* Windows and their peak fits on the device (the per-pair body of get_image_shift,
 * imageprocess.py:103-157: arg-max, 5x5 cut-out, Gaussian peak fit with the reference's start
 * values and bounds).  peak_records (n_pairs, 32) float64 per pair: [0] status (0 fitted,
 * 1 not settled -> re-fit the 5x5 window [5..29] on the host with the reference's curve_fit,
 * 2 cut-out empty / not square -> shift (0, 0), 3 square but not 5x5 -> host needs the window),
 * [1] y_max, [2] x_max, [3] xc, [4] yc.  `windows` is optional (NULL: 256 bytes per pair come
 * back instead of H x W floats). */
int pb_undrift_peaks_pairs(int n_seg, const long long* seg_start, const float* x, const float* y,
                           const float* lpx, const float* lpy, int Y, int X, double min_blur_width,
                           int Y0, int X0, int H, int W, int n_pairs, const int* pair_i,
                           const int* pair_j, double* peak_records, double* sums, float* windows);

#ifdef __cplusplus
}
#endif
#endif /* PICASSO_B200_H */
