/*
 * fava_b200.h — C ABI of libfava_b200.so, the B200 (sm_100a) implementation of FAVA's
 * grid-statistics hot path.
 *
 * The reference (ebrooker/FAVA) is pure Python and has no FFI of its own; the boundary a
 * maintainer would bind is therefore the set of NumPy loops this library replaces.  Each entry
 * point below cites the reference code it stands in for (paths relative to the reference root).
 * INTEGRATION.md shows the ctypes stub that binds them from the reference's mesh classes.
 *
 * Conventions
 *   - plain C types only: pointers, sizes, ints, doubles.  No torch / C++ types cross the ABI.
 *   - every pointer named d_* is a DEVICE pointer owned by the caller; h_* is a HOST pointer.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls enqueue
 *     work on that stream and return; results are valid after the stream is synchronised
 *     (fava_stream_sync or the caller's own event).
 *   - all functions return FAVA_OK (0) or a negative FAVA_E* code and never throw; the message
 *     for the calling thread's last failure is fava_last_error().
 *   - field arrays use the FLASH *file* layout: [z][y][x] (x fastest) for a uniform dataset,
 *     [block][z][y][x] for a block dataset.  The reference's in-memory layout [x][y][z]
 *     (fava/mesh/FLASH/_flash.py:314-335, a strided transpose on load) is never materialised.
 *   - dtype: FAVA_F32 (plt files) or FAVA_F64 (chk files); f32 is widened to f64 in registers,
 *     bit-identical to `.astype(np.float64)` (_flash.py:333).  All arithmetic is fp64.
 *   - reductions are deterministic: fixed-order two-level accumulation, no floating-point atomics.
 *   - a fava_ctx serves ONE compute stream at a time (its scratch buffers, cached tables and staging ring are not
 *     fenced between streams); the one supported concurrent pair is fava_fft_y_scatter on a side stream beside the
 *     plane-moment kernels / fava_ke_transform_z on another (they share no scratch).  Use one context per device.
 */
#ifndef FAVA_B200_H
#define FAVA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FAVA_ABI_VERSION 4

/* status codes */
#define FAVA_OK 0
#define FAVA_EINVAL (-1)  /* bad argument (shape, dtype, axis, null pointer) */
#define FAVA_ECUDA (-2)   /* CUDA runtime / cuFFT error; see fava_last_error() */
#define FAVA_ENOMEM (-3)  /* workspace allocation failed */
#define FAVA_EIO (-4)     /* staging: open/pread failed or short read */
#define FAVA_ENODEV (-5)  /* no usable sm_100 device */

/* storage dtypes of field arrays */
#define FAVA_F32 0
#define FAVA_F64 1

/* number of moment rows produced per bin by the plane-moment kernels */
#define FAVA_NMOM 14
/*
 * Moment rows (all vf-weighted where a weight applies), with d_i = u_i - c_i (c = pivot):
 *   0        S0      = sum rho
 *   1..3     Sd_i    = sum d_i                  (i = x,y,z)
 *   4..6     Srd_i   = sum rho d_i
 *   7..12    Srdd_ij = sum rho d_i d_j          (xx,xy,xz,yy,yz,zz)
 *   13       W       = sum of weights (cells x vol_frac)
 */

typedef struct fava_ctx fava_ctx;

/* ---- context ------------------------------------------------------------------------------- */

/* Create a context bound to CUDA device `device` (the library owns workspaces, cuFFT plans and
 * the pinned staging ring inside it).  Replaces the reference's FAVA_MPI singleton +
 * shared-memory windows (fava/util/_mpi.py:7-80). */
int fava_init(int device, fava_ctx** out);
int fava_shutdown(fava_ctx* ctx);
int fava_stream_sync(fava_ctx* ctx, void* stream);
const char* fava_last_error(void);
int fava_abi_version(void);
/* Number of kernels this library has launched since load (bench.py's gpu_launches). */
int64_t fava_launch_count(void);

/* ---- Reynolds / Favre plane statistics (reference: FLASH.reynolds_stress, _flash.py:1506-1611) */

/* Per-bin pivots c_i[n] = u_i at the first cell of plane n (dense array).  d_pivots: [3][nbins]. */
int fava_plane_pivots(fava_ctx* ctx, const void* d_ux, const void* d_uy, const void* d_uz, int dtype,
                      int64_t nz, int64_t ny, int64_t nx, int axis, double* d_pivots, void* stream);

/* One streaming pass over rho,ux,uy,uz [nz][ny][nx]: the FAVA_NMOM pivoted plane moments for
 * every plane normal to `axis` (0 = x, the fastest index; 1 = y; 2 = z), unweighted.
 * Replaces both hot loops of the reference (_flash.py:1564-1577 and :1584-1604).
 * d_moments: [FAVA_NMOM][nbins], nbins = (nx,ny,nz)[axis].  If `accumulate` != 0 the result is
 * added to d_moments (chunk-streamed slabs; axis-2 callers pass bin_offset instead).
 * Row 13 (W) receives the cell count of each plane. */
int fava_plane_moments(fava_ctx* ctx, const void* d_rho, const void* d_ux, const void* d_uy,
                       const void* d_uz, int dtype, int64_t nz, int64_t ny, int64_t nx, int axis,
                       const double* d_pivots, double* d_moments, int accumulate, void* stream);

/* Axis x AND axis z from ONE pass over the data: the column kernel runs with one chunk per z-plane; its per-plane
 * column partials are summed over z for the x-bins and, re-expressed about the plane pivots (exact algebra),
 * over x for the z-bins.  Saves one of the three passes of an x/y/z profile set.
 * d_piv_x [3][nx], d_piv_z [3][nz] (fava_plane_pivots with axis 0 / 2); d_mom_x [FAVA_NMOM][nx], d_mom_z [FAVA_NMOM][nz]. */
int fava_plane_moments_xz(fava_ctx* ctx, const void* d_rho, const void* d_ux, const void* d_uy, const void* d_uz,
                          int dtype, int64_t nz, int64_t ny, int64_t nx, const double* d_piv_x,
                          const double* d_piv_z, double* d_mom_x, double* d_mom_z, void* stream);
/* All three profile sets from ONE pass over the fields (32 B/cell fp64 instead of 64): moments for axis x, y and z
 * with the pivots of fava_plane_pivots (axis 0 / 1 / 2).  Needs nx % 256 == 0 and ny % 8 == 0
 * (fava_plane_moments_xyz_supported); other shapes use fava_plane_moments_xz + fava_plane_moments. */
int fava_plane_moments_xyz_supported(int64_t nz, int64_t ny, int64_t nx);
int fava_plane_moments_xyz(fava_ctx* ctx, const void* d_rho, const void* d_ux, const void* d_uy, const void* d_uz,
                           int dtype, int64_t nz, int64_t ny, int64_t nx, const double* d_piv_x, const double* d_piv_y,
                           const double* d_piv_z, double* d_mom_x, double* d_mom_y, double* d_mom_z, void* stream);

/* Block-list front end for FLASH block datasets [nblocks][nzb][nyb][nxb] (AMR or multi-block
 * uniform plt files).  For leaf l of the table: planes i=0..nrb-1 of block blk[l] normal to `axis`
 * contribute weight vf[l] to fine bins [ilo[l]+i*scale[l], ilo[l]+(i+1)*scale[l])
 * (_flash.py:1559-1577, :1594-1604).  The table is built on the host exactly as the reference
 * does (get_blocklist :803-822, vol_fracs :1559-1562, ilo :1566-1567, lref_n :1565).
 * Outputs d_moments [FAVA_NMOM][nbins] and d_pivots [3][nbins] (pivot chosen per bin by the
 * library: first cell of the first contributing block plane). */
typedef struct fava_leaf_desc {
    int64_t block;   /* index of the source block in the dataset */
    int64_t ilo;     /* first fine bin covered by plane 0 */
    int32_t scale;   /* lref_n = 2^(lmax - level): fine bins per block plane */
    int32_t pad_;
    double vol_frac; /* vf_b */
} fava_leaf_desc;

int fava_plane_moments_blocks(fava_ctx* ctx, const void* d_rho, const void* d_ux, const void* d_uy,
                              const void* d_uz, int dtype, int64_t nzb, int64_t nyb, int64_t nxb,
                              int axis, const fava_leaf_desc* h_leaves, int64_t nleaf, int64_t nbins,
                              double* d_moments, double* d_pivots, void* stream);

/* Same, for callers whose leaf tables are immutable objects: `table_uid` != 0 names the table, and two calls with
 * equal uid, nleaf and nbins are taken to pass identical tables - the device-side tables cached from the previous
 * call are reused without comparing the 32 * nleaf bytes again (0.7 ms of host time at 262144 leaves, more than
 * the kernels).  table_uid = 0: compare, as fava_plane_moments_blocks does. */
int fava_plane_moments_blocks_uid(fava_ctx* ctx, const void* d_rho, const void* d_ux, const void* d_uy,
                                  const void* d_uz, int dtype, int64_t nzb, int64_t nyb, int64_t nxb,
                                  int axis, const fava_leaf_desc* h_leaves, int64_t nleaf, uint64_t table_uid,
                                  int64_t nbins, double* d_moments, double* d_pivots, void* stream);

/* Re-express moments taken about pivots c_old about c_new (exact algebra; used before summing
 * partial moments from different ranks / slabs whose pivots differ). In place on d_moments. */
int fava_moments_repivot(fava_ctx* ctx, double* d_moments, const double* d_piv_old,
                         const double* d_piv_new, int64_t nbins, void* stream);

/* Moments -> profiles.  weight = vol_frac applied to unweighted (dense) moments, 1.0 for the
 * block-list front end (already weighted); layer_volume as _flash.py:1526-1542.
 *   d_means  [4][nbins]: dens, velx, vely, velz volume means            (_flash.py:1579-1582)
 *   d_rey    [6][nbins]: <rho u'_i u'_j>  xx,xy,xz,yy,yz,zz              (_flash.py:1597-1609)
 *   d_fmeans [3][nbins]: Favre means  u~_i = <rho u_i>/<rho>             (extension)
 *   d_favre  [6][nbins]: <rho u''_i u''_j>                               (extension)
 * Any output pointer may be NULL. */
int fava_moments_finalize(fava_ctx* ctx, const double* d_moments, const double* d_pivots,
                          int64_t nbins, double weight, double layer_volume, double* d_means,
                          double* d_rey, double* d_fmeans, double* d_favre, void* stream);

/* Single-moment variant: vf-weighted plane integral of one field (reference: slice_integral,
 * _flash.py:1451-1504).  Dense array; d_out [nbins] = sum over plane (unweighted). */
int fava_plane_sum(fava_ctx* ctx, const void* d_field, int dtype, int64_t nz, int64_t ny, int64_t nx,
                   int axis, double* d_out, void* stream);
/* Block-list variant: d_out [nbins] = sum over leaves of vol_frac * plane sums, scattered to the fine
 * bins like fava_plane_moments_blocks (_flash.py:1488-1498). */
int fava_plane_sum_blocks(fava_ctx* ctx, const void* d_field, int dtype, int64_t nzb, int64_t nyb, int64_t nxb,
                          int axis, const fava_leaf_desc* h_leaves, int64_t nleaf, int64_t nbins, double* d_out,
                          void* stream);
int fava_plane_sum_blocks_uid(fava_ctx* ctx, const void* d_field, int dtype, int64_t nzb, int64_t nyb, int64_t nxb,
                              int axis, const fava_leaf_desc* h_leaves, int64_t nleaf, uint64_t table_uid,
                              int64_t nbins, double* d_out, void* stream);

/* ---- AMR -> uniform prolongation (reference: FLASH.from_amr gather, _flash.py:1262-1321) ----- */

typedef struct fava_prolong_leaf {
    int64_t block;    /* source block index */
    int32_t off[3];   /* fine-cell corner of the block minus subdomain corner: x,y,z (may be <0) */
    int32_t scale;    /* 2^(L - level) */
} fava_prolong_leaf;

/* Piecewise-constant injection of the selected leaves into a uniform [NZ][NY][NX] fp64 array.
 * Later table entries win where leaves overlap (the reference's dict overwrite order,
 * _flash.py:1305); cells no leaf covers are 0.0 (_flash.py:1258). */
int fava_prolong(fava_ctx* ctx, const void* d_blocks, int dtype, int64_t nzb, int64_t nyb,
                 int64_t nxb, const fava_prolong_leaf* h_leaves, int64_t nleaf, int64_t NZ,
                 int64_t NY, int64_t NX, double* d_out, void* stream);

/* ---- kinetic-energy spectrum (reference: FlashUniform.kinetic_energy_spectra,
 *      fava/mesh/FLASH/FlashUniform.py:229-304) -------------------------------------------------- */

/* Whole pipeline on one GPU for a cubic N^3 grid: w_n = sqrt(rho) u_n, 3-D FFT, |u^|^2 / longitudinal
 * projection / shell binning, shell means x 4 pi k^2.  Power-of-two N in [256, 2048] take the hand-written
 * transform (csrc/fft.cu); any other even N uses cuFFT D2Z/Z2Z - the one library call on this path.
 * Outputs are HOST arrays of nbins = N/2 - 1 doubles (keys k,total,longitudinal,transverse of the
 * reference's dict). */
int fava_ke_spectrum(fava_ctx* ctx, const void* d_rho, const void* d_ux, const void* d_uy,
                     const void* d_uz, int dtype, int64_t n, double* h_k, double* h_total,
                     double* h_long, double* h_trans, void* stream);

/* Building blocks of the same pipeline, exposed for the slab-decomposed (multi-GPU) and the streamed
 * (host-resident snapshot) drivers.  Spectral arrays are Hermitian halves, complex [..][..][pitch] with
 * pitch = fava_spectral_pitch(n) complex numbers per kx row: n/2 on the hand-written path (the Nyquist column
 * kx = n/2 lies beyond the last bin edge n/2 - 1.5, FlashUniform.py:273-276, and is not stored, which keeps
 * every row a whole number of 128-byte lines), n/2 + 1 on the cuFFT path. */
int64_t fava_spectral_pitch(int64_t n);
int fava_fft_native_supported(int64_t n); /* 1: n is a power of two in [256, 2048] */

/* Stage 1 of the transform of a z-slab [nz_local][n][n] of rho,ux,uy,uz (FlashUniform.py:266-268): on the
 * hand-written path the weighting w_c = sqrt(rho) u_c FUSED with the x transform (the weighted real fields are
 * never written), on the cuFFT path the weighting alone (real rows padded to 2 pitch doubles).  d_wx/wy/wz:
 * one buffer of 16 nz_local n pitch bytes per component. */
int fava_ke_transform_x(fava_ctx* ctx, const void* d_rho, const void* d_ux, const void* d_uy, const void* d_uz,
                        int dtype, int64_t nz_local, int64_t n, double* d_wx, double* d_wy, double* d_wz,
                        void* stream);
/* Stage 2, per component, in place: completes the 2-D transform of the slab -> complex [nz_local][ky][kx]. */
int fava_ke_transform_y(fava_ctx* ctx, double* d_w, int64_t nz_local, int64_t n, void* stream);
/* Stage 3, per component, in place: transform along z of complex [n (z)][ny_local][pitch].  Row jl holds global
 * ky index d_ky_of_local[jl] (NULL = this GPU holds every ky in order; -1 = padding row).  The hand-written pass
 * skips columns outside the spectral disc and output rows outside the sphere: elements with
 * kx^2 + ky^2 + kz^2 > (n/2 - 1.5)^2 are unspecified afterwards (no bin reads them). */
int fava_ke_transform_z(fava_ctx* ctx, double* d_w, int64_t n, int64_t ny_local, const int32_t* d_ky_of_local,
                        void* stream);

/* The kernels behind the stages (also used directly by the tests). */

/* K4: w_n = sqrt(rho) * u_n for n = x,y,z in one pass.  Inputs are `nrows` rows of `nx` cells; outputs are
 * real fp64 rows of `pitch` doubles (pitch >= nx, even). */
int fava_ke_weight3(fava_ctx* ctx, const void* d_rho, const void* d_ux, const void* d_uy, const void* d_uz,
                    int dtype, int64_t nrows, int64_t nx, int64_t pitch, double* d_wx, double* d_wy,
                    double* d_wz, void* stream);
/* x pass fused with the weighting (nx a power of two in [256, 2048], nrows a multiple of 8): reads `nrows`
 * rows of nx cells of rho,ux,uy,uz once and writes, per component, kx = 0..nx/2-1 of the spectrum of
 * sqrt(rho)*u along x into complex rows of `pitch` elements. */
int fava_fft_x_weight3(fava_ctx* ctx, const void* d_rho, const void* d_ux, const void* d_uy, const void* d_uz,
                       int dtype, int64_t nrows, int64_t nx, int64_t pitch, double* d_fx, double* d_fy,
                       double* d_fz, void* stream);
/* In-place complex FFT of length n along dimension `line_dim` (1: d1, 2: d2) of complex [d2][d1][pitch], for
 * the columns kx < ncols (a multiple of 8192/n).  prune_mode 0: every output; 1 (y pass): output rows with
 * ky^2 + kx0^2 > (n/2-1.5)^2 are not written (kx0 = first column of the 8192/n-wide column tile); 2 (z pass,
 * batch rows = ky rows with global index d_ky_of_batch[b], NULL = identity): tiles with ky^2 + kx0^2 beyond
 * the disc are skipped and output rows beyond the sphere are not written. */
int fava_fft_cols(fava_ctx* ctx, double* d_data, int64_t n, int64_t pitch, int64_t ncols, int64_t d1, int64_t d2,
                  int line_dim, int prune_mode, const int32_t* d_ky_of_batch, void* stream);
/* Stage 2 on several GPUs = y pass FUSED with the slab -> ky-pencil exchange (hand-written path only): the planes
 * [z_offset, z_offset + nz_chunk) of this rank's slab, given at d_data as complex [nz_chunk][n][n/2] (x-transformed),
 * are transformed along y and every output row ky is stored straight into the receive buffer of the rank that owns
 * it, d_peer_recv[d_owner_of_ky[ky]] (peer-mapped device memory, or local when that is my_rank), at
 * [my_rank*nz_local + z][d_row_of_ky[ky]][kx] of that rank's complex [n][nyl][n/2] array - the layout stage 3
 * consumes.  Rows nobody owns (d_owner_of_ky = -1: the Nyquist row) and rows outside the spectral disc are not sent.
 * max_ctas > 0 limits the persistent grid (NVLink needs fewer SMs than HBM; the rest run other kernels).  The stores are
 * complete on every rank once all ranks have passed a stream-ordered collective issued after this call. */
int fava_fft_y_scatter(fava_ctx* ctx, double* d_data, int64_t n, int64_t nz_chunk, double* const* d_peer_recv,
                       const int32_t* d_owner_of_ky, const int32_t* d_row_of_ky, int my_rank, int64_t nz_local,
                       int64_t nyl, int64_t z_offset, int max_ctas, void* stream);
/* Slab -> ky-pencil exchange as a kernel of its own (cuFFT path; the hand-written path fuses it into stage 2,
 * fava_fft_y_scatter): rank `my_rank` holds complex [nz_local][n][pitch] after stage 2; spectral space is distributed over ky in +-ky symmetric sets (so the transposed operand of
 * the longitudinal projection stays rank-local).  For every destination rank r the kernel gathers the ky
 * rows owned by r (d_ky_of_dest: [nranks][nyl] global ky indices, -1 = padding) and writes them straight
 * into r's receive buffer d_peer_recv[r] (peer-mapped device memory, or the local buffer when r ==
 * my_rank) at [my_rank*nz_local + z][row][kx] of a complex [n][nyl][pitch] array: no staging copy, the
 * NVLink stores overlap the gather (TMA bulk copies global -> shared -> peer global).  Columns outside the
 * spectral disc are not sent. */
int fava_a2a_pack(fava_ctx* ctx, const double* d_in, double* const* d_peer_recv, const int32_t* d_ky_of_dest,
                  int my_rank, int nranks, int64_t nz_local, int64_t n, int64_t nyl, void* stream);
/* Shell binning of one spectral sub-volume complex [n (kz)][ny_local][pitch] x 3 components of an n^3
 * transform scaled by `norm` (1/n^3, norm="forward"), row pitch fava_spectral_pitch(n).  Row jl holds global ky index
 * d_ky_of_local[jl] (NULL = this GPU holds every ky in order; -1 = padding row).  One streaming pass over the elements
 * inside the spectral sphere: the reference's `.T` projection (FlashUniform.py:281) is evaluated at every point from
 * the values stored AT that point, | kz u^_x + ky u^_y + kx u^_z |^2 / |k|^2 - the shell sums are those of the
 * reference's transposed form (csrc/spectrum.cu), so the ky rows may be distributed over the ranks in any way.
 * d_sums: [3][n/2-1] = weighted sums of total, longitudinal, and the point counts (FlashUniform.py:273-293). */
int fava_spectrum_bin(fava_ctx* ctx, const double* d_fx, const double* d_fy, const double* d_fz, int64_t n,
                      int64_t ny_local, const int32_t* d_ky_of_local, double norm, double* d_sums, void* stream);
/* Shell sums -> spectra (host outputs, nbins = n/2-1 each): mean x 4 pi k^2 (FlashUniform.py:286-302). */
int fava_spectrum_finalize(fava_ctx* ctx, const double* d_sums, int64_t n, double* h_k, double* h_total,
                           double* h_long, double* h_trans, void* stream);

/* ---- uniform-grid analyses next to the spectrum ------------------------------------- */

/* Box counting of an iso-contour (reference: FlashUniform.fractal_dimension, FlashUniform.py:85-227).
 * d_field holds planes [zf0, zf1) of a [nz][ny][nx] field; the call flags the contour cells of planes [z0, z1)
 * (edge marking, FlashUniform.py:114-177: the rule int((c - val) / (nb - val)) == 0 decides between the low
 * cell and its neighbour) and adds to d_counts[level], level = 0..5, the number of boxes of edge 2^level cells
 * that hold a flag (:179-208); d_coarse [ceil(nz/32)][ceil(ny/32)][ceil(nx/32)] receives the occupancy byte of
 * every 32^3 tile it visits.  [z0, z1) must be aligned to 32 planes (z1 may be nz) and the buffer must include
 * one halo plane on each inner side.  The caller zero-fills d_counts (FAVA_FRACTAL_MAXLEVELS entries) and
 * d_coarse beforehand; several ranks cover disjoint plane ranges and sum both arrays. */
#define FAVA_FRACTAL_MAXLEVELS 32
int fava_fractal_tiles(fava_ctx* ctx, const void* d_field, int dtype, int64_t nz, int64_t ny, int64_t nx,
                       int64_t zf0, int64_t zf1, int64_t z0, int64_t z1, double contour, uint64_t* d_counts,
                       uint8_t* d_coarse, void* stream);
/* Levels 6 .. nlevels-1 (box edge 2^(level-5) tiles) from the complete tile-occupancy grid of an [nz][ny][nx]
 * field, added to d_counts[level]. */
int fava_fractal_coarse(fava_ctx* ctx, const uint8_t* d_coarse, int64_t nz, int64_t ny, int64_t nx, int nlevels,
                        uint64_t* d_counts, void* stream);

/* Structure functions (reference: FlashUniform.structure_functions, FlashUniform.py:306-445).
 * fava_sf_gather: d_points [npoints][3] (x,y,z) -> cell index floor((p - lo) / cell) per axis (:400-406) ->
 * d_vel [npoints][3] = (velx, vely, velz) of that cell (:408-412) when its plane lies in the held range
 * [zf0, zf1) of the [nz][ny][nx] fields, 0.0 otherwise (ranks sum their parts).  *d_err is set to 1 if a point
 * falls outside the grid (the reference raises IndexError there); the caller zeroes it. */
int fava_sf_gather(fava_ctx* ctx, const double* d_points, int64_t npoints, const void* d_ux, const void* d_uy,
                   const void* d_uz, int dtype, int64_t nz, int64_t ny, int64_t nx, int64_t zf0, int64_t zf1,
                   const double* h_lo, const double* h_cell, double* d_vel, int* d_err, void* stream);
/* fava_sf_moments: for each of nsep separations, over its npoints pairs (p1, p2 coordinates, v1, v2 velocities,
 * all [nsep][npoints][3]): r^ = (p2 - p1)/|p2 - p1| (or (1,0,0) if anisotropic), dl = |dv . r^|,
 * dt = |dv - dl r^|; d_out [2][nsep] = mean dl^order, mean dt^order (:417-436). */
int fava_sf_moments(fava_ctx* ctx, const double* d_p1, const double* d_p2, const double* d_v1, const double* d_v2,
                    int64_t nsep, int64_t npoints, int order, int anisotropic, double* d_out, void* stream);

/* ---- HDF5 block staging (reference: _read_variable_data, _flash.py:306-341) ----------------- */

/* Stream `nbytes` at `file_offset` of a contiguous HDF5 dataset (offset from the h5lite index)
 * into device memory through the context's pinned ring: pread by reader threads, one
 * cudaMemcpyAsync per chunk on `stream`.  Returns after the last copy is enqueued. */
int fava_stage_h2d(fava_ctx* ctx, const char* path, int64_t file_offset, int64_t nbytes,
                   void* d_dst, void* stream);
/* Same path for a caller-owned host buffer (pageable or pinned). */
int fava_stage_host_h2d(fava_ctx* ctx, const void* h_src, int64_t nbytes, void* d_dst,
                        void* stream);

/* ---- context-owned device buffers and CUDA IPC helpers for peer-mapped exchange buffers
 *      (one process per GPU) ------------------------------------------------------------------------ */
/* Grow-only device buffer owned by the context (a plain cudaMalloc allocation, so that it can be
 * exported with fava_ipc_export); zero-filled when (re)allocated.  Slots 8..15 are free for callers. */
#define FAVA_WS_USER0 8
#define FAVA_WS_NSLOTS 20
int fava_workspace(fava_ctx* ctx, int slot, int64_t bytes, void** d_ptr_out);
int fava_ipc_export(void* d_ptr, unsigned char handle_out[64]);
int fava_ipc_open(const unsigned char handle[64], void** d_ptr_out);
int fava_ipc_close(void* d_ptr);

#ifdef __cplusplus
}
#endif
#endif /* FAVA_B200_H */
