/* libfsg — C-ABI of the B200-native FetalSynthGen per-sample generation path.
 *
 * Boundary: every entry point replaces the torch-eager body of one stage of the reference's
 * FetalSynthGen.sample() (fetalsyngen/generator/model.py:94-276).  The reference has no FFI
 * for the base stages (they are torch calls) and a pybind11 one for the motion artifact
 * (svort/slice_acquisition/slice_acq_cuda.cpp:156-161); a maintainer binds this library with
 * ctypes.  Signatures carry plain pointers, sizes and an opaque stream
 * handle only — no torch types.
 *
 * Conventions
 *  - Volumes are C-contiguous [x][y][z] (z fastest), float32 images, uint8 label maps.
 *  - All data pointers are DEVICE pointers unless the name ends in _host.
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *  - Every call processes a batch of `njobs` independent volumes that share one shape
 *    (njobs <= FSG_MAX_JOBS); job structs are HOST memory, copied by value into the launch.
 *  - Return value: 0 = ok, non-zero = error; fsg_last_error() gives the message
 *    (thread-local).  Nothing allocates device memory inside a stage call.
 *  - No CPU fallback exists: without a CUDA device every launch returns an error.
 */
#ifndef FSG_H
#define FSG_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FSG_VERSION 102
#define FSG_MAX_JOBS 16
#define FSG_MAX_TAPS 127

/* One entry of a 1-D linear sampling table (output index -> source cell).
 * Built on the host with the reference's own expressions (torch.arange float32 positions,
 * utils/generation.py:318-361; numpy float64 positions cast to float32, synthseg.py:84-102)
 * because their float32 rounding is implementation-defined.  f < 0 marks a position outside
 * (0, S-1] for which linear sampling yields 0 (utils/generation.py:229-235). */
typedef struct fsg_tab {
  int16_t f;  /* floor index */
  int16_t c;  /* ceil index, clamped to S-1 */
  float wc;   /* weight of the ceil sample; floor weight is 1 - wc */
} fsg_tab;

/* Counter-based RNG stream: Philox4x32-10, key = seed, counter = (block, stage, sample). */
typedef struct fsg_rng {
  uint64_t seed;
  uint64_t sample;
  uint32_t stage;
  uint32_t _pad;
} fsg_rng;

/* Block-linear intensity volume: a layered 2-D CUDA array (layer = x, row = y, column = z; float32) with a
 * surface object, through which fsg_gmm writes it, and a point-sampled texture object, through which
 * fsg_warp's fast path fetches the 2x2 (y, z) footprint of a trilinear sample with ONE texture gather (tld4)
 * per x layer: two texture instructions per voxel instead of eight global loads, on the texture unit's
 * tiled addressing instead of the load/store unit's 128-byte lines (the gather of a rotated row crosses
 * many lines; measured r02: 0.92 -> 0.66 ms per 8 volumes of 256^3).  The texel values are the same float32 bits,
 * so results are identical to the linear path.  Handles are plain 64-bit integers (cudaArray_t,
 * cudaTextureObject_t, cudaSurfaceObject_t). */
typedef struct fsg_texvol {
  uint64_t array;
  uint64_t tex;
  uint64_t surf;
  int32_t nx, ny, nz, _pad;
} fsg_texvol;
int fsg_texvol_create(int nx, int ny, int nz, fsg_texvol* out);
int fsg_texvol_destroy(fsg_texvol* v);
/* linear [nx][ny][nz] float32 device volume <-> the array (to_linear != 0: array -> linear); tests and debugging */
int fsg_texvol_copy(const fsg_texvol* v, float* linear_dev, int to_linear, void* stream);

/* K1 — GMM intensity synthesis. Replaces ImageFromSeeds.sample_intensities' per-voxel part
 * (generator/intensity/rand_gmm.py:146-149) fused with the seed sum of load_seeds
 * (rand_gmm.py:90-97): L = sum(seed[m]); I = max(0, mus[L] + sigmas[L] * N). */
typedef struct fsg_gmm_job {
  const int8_t* seed[4]; /* 1..4 label volumes summed per voxel; unused entries NULL */
  const float* mus;      /* [nlabels] */
  const float* sigmas;   /* [nlabels] */
  const float* noise;    /* [nvox] standard normal draws to inject, or NULL -> Philox */
  float* out;            /* [nvox], or NULL when out_pairs is used */
  uint8_t* labels_out;   /* optional [nvox] summed labels, or NULL */
  fsg_rng rng;
  int32_t nlabels;
  int32_t row_len;       /* z extent of a row (out_pairs only): pairs do not cross row ends */
  uint32_t* out_pairs;   /* optional [nvox]: 16.7 fixed point pairs (I[v] | I[v+1] << 16), I * 128 rounded;
                          * the gather format of fsg_warp's fast path (one 32-bit load brings both z corners;
                          * quantisation <= 1/256 intensity unit, ~1e-5 of the intensity range) */
  int32_t pairs_float;   /* != 0: out_pairs is a float2 volume [nvox][2] = (I[v], I[v+1]) instead — lossless,
                          * one 8-byte gather per (x, y) row in fsg_warp; twice the bytes written here */
  int32_t word_bytes;    /* packed mode (words != NULL): 2 or 4 */
  /* packed mode: the sample's labels come straight from a subject's bit-packed seed words (see fsg_unpack_job:
   * L = meta ? 10 * meta + ((word >> shift[meta-1]) & mask[meta-1]) : 0) instead of the sum of seed volumes, so
   * the dataset-cache path reads 2 bytes per voxel and needs no fsg_unpack_seeds pass.  seed[] is ignored. */
  const void* words;     /* [nvox] uint16 / uint32 (device), 16-byte aligned, or NULL */
  int32_t shift[4];
  int32_t mask[4];
  uint64_t out_surf;     /* fsg_texvol.surf: write the intensities into the block-linear volume instead of `out`
                          * (row_len = nz, surf_ny = ny; nvox = nx * ny * nz, nz % 4 == 0), or 0 */
  int32_t surf_ny;
  int32_t _pad2;
} fsg_gmm_job;
int fsg_gmm(const fsg_gmm_job* jobs_host, int njobs, int64_t nvox, void* stream);

/* K2 — fused spatial deformation. Replaces SpatialDeformation.deform
 * (generator/deformation/affine_nonrigid.py:86-120, 164-193, 299-366) + myzoom_torch of the
 * control grid (utils/generation.py:310-397) + fast_3D_interp_torch linear/nearest
 * (utils/generation.py:204-288), with RandGamma (synthseg.py:262-275) and RandBiasField
 * (synthseg.py:157-188) as optional epilogues on the image.
 * mode 0: identity sampling (deformation gate off) — only flip + epilogues are applied.
 * mode 1: coordinates = A*(grid - center + F) + c2, clamped to [0,S-1], minus shift. */
typedef struct fsg_warp_job {
  const float* src_img;   /* [S] or NULL */
  const uint32_t* src_pairs; /* fsg_gmm's out_pairs volume instead of src_img (fast path only), or NULL */
  const uint8_t* src_seg; /* [S] or NULL */
  const float* src_img2;  /* optional second image (load_image=True), or NULL */
  float* dst_img;
  uint8_t* dst_seg;
  float* dst_img2;
  const float* fsmall;    /* [fs0][fs1][fs2][3] control-grid displacements, or NULL */
  const fsg_tab* ftab[3]; /* zoom tables of the control grid, lengths sx, sy, sz */
  const float* bf_low;    /* [bs0][bs1][bs2] log-bias control grid, or NULL */
  const fsg_tab* btab[3]; /* zoom tables of the bias grid */
  const float* shift;     /* [3] floor(min coord) per axis (written by fsg_warp_shift) */
  float A[9];
  float c2[3];
  float center[3];
  float gamma;            /* applied when has_gamma */
  int32_t fs[3];
  int32_t bs[3];
  int32_t mode;
  int32_t flip;
  int32_t has_gamma;
  int32_t pairs_float;    /* src_pairs holds fsg_gmm's float2 pairs (pairs_float there) instead of fixed point */
  uint64_t src_tex;       /* fsg_texvol.tex of the source image instead of src_img (fast path only), or 0 */
} fsg_warp_job;
/* Pre-pass: writes floor(min over the volume of the clamped coordinate) per axis into
 * job.shift (affine_nonrigid.py:350-358).  No volume traffic; coordinates are recomputed. */
int fsg_warp_shift(const fsg_warp_job* jobs_host, int njobs, int sx, int sy, int sz, void* stream);
int fsg_warp(const fsg_warp_job* jobs_host, int njobs, int sx, int sy, int sz, void* stream);
/* Debug/parity aid: writes the three coordinate volumes (after clamp and shift). */
int fsg_warp_coords(const fsg_warp_job* job_host, int sx, int sy, int sz, float* xx, float* yy, float* zz, void* stream);

/* K4a — separable zero-padded Gaussian blur. Replaces gaussian_blur_3d
 * (utils/generation.py:84-110); taps come from make_gaussian_kernel on the host. */
typedef struct fsg_blur_job {
  const float* src;
  float* dst;
  float* tmp;            /* scratch volume (same size) */
  const float* taps[3];  /* per axis, or NULL / ntaps 0 to skip the axis */
  int32_t ntaps[3];
  int32_t _pad;
} fsg_blur_job;
int fsg_blur3d(const fsg_blur_job* jobs_host, int njobs, int sx, int sy, int sz, void* stream);

/* K4ab — fused resolution simulation: separable banded resampling out = (Rx (x) Ry (x) Rz) in,
 * with the RandNoise epilogue.  Replaces gaussian_blur_3d followed by the trilinear
 * down-sampling of RandResample.__call__ (synthseg.py:63-107; utils/generation.py:84-110,
 * 227-285) and RandNoise (synthseg.py:217-235): per axis the Gaussian taps and the two linear
 * interpolation weights compose into one banded matrix whose rows the host supplies
 * (window start q0 and `width` weights per output; zero weights = zero padding / positions the
 * reference's sampler maps to 0).  With identity positions it is a plain separable blur.
 * Requirements: q0 non-decreasing, 0 <= q0, q0 + width <= n_in.
 * Passes run x, y, z: x  src -> tmp1 [n_out0][sy][sz],  y  tmp1 -> tmp2 [n_out0][n_out1][sz],  z  tmp2 -> dst
 * (+ noise).  When the z axis is also given uncomposed (`pos` / `taps`, <= 13 taps) and sz <= 256, the last pass
 * blurs and samples rows in registers instead of applying the composed windows from shared memory; results agree
 * to the float tolerance.  Philox noise: block = row * ceil(n_out2 / 4) + K / 4 of output row (I, J), component K % 4.
 * cap_* are the capacities of the buffers in floats: the call fails instead of writing past them (an axis
 * may be up-sampled, n_out > n_in, when the simulated spacing is finer than the input resolution). */
typedef struct fsg_sepaxis {
  const int16_t* q0; /* [n_out] first source index of each output's window */
  const float* w;    /* [n_out][width] window weights */
  int32_t n_out;
  int32_t width;
  /* the same axis before composition (what fsg_sep_compose was given); used by the fused schedule for z */
  const fsg_tab* pos; /* [n_out] sampling table, NULL = identity positions */
  const float* taps;  /* [ntaps] Gaussian taps, NULL = no blur */
  int32_t ntaps;
  int32_t _pad;
} fsg_sepaxis;
typedef struct fsg_sepconv_job {
  const float* src;  /* [sx][sy][sz] */
  float* dst;        /* [n_out0][n_out1][n_out2] */
  float* tmp1;
  float* tmp2;
  fsg_sepaxis ax[3];
  const float* noise; /* [n_out0*n_out1*n_out2] injected draws or NULL -> Philox (when has_noise) */
  fsg_rng rng;
  float noise_std;
  int32_t has_noise;
  int64_t cap_dst, cap_tmp1, cap_tmp2; /* floats available behind dst / tmp1 / tmp2 */
} fsg_sepconv_job;
int fsg_sepconv(const fsg_sepconv_job* jobs_host, int njobs, int sx, int sy, int sz, void* stream);

/* Builds one axis table of fsg_sepconv on the device: row I = w_f*taps(.-f_I) + w_c*taps(.-c_I)
 * over the window [q0, q0+width), width = min(n_in, ntaps + (pos != NULL)).
 * `pos` is the 1-D sampling table of the coarse grid (f < 0: the output is 0), NULL = identity
 * positions; `taps` are the zero-padded Gaussian taps (make_gaussian_kernel,
 * utils/generation.py:74-81), NULL = no blur on this axis.  Up to 3*FSG_MAX_JOBS tables per call. */
typedef struct fsg_sepcompose_job {
  const fsg_tab* pos;
  const float* taps;
  int16_t* q0_out; /* [n_out] */
  float* w_out;    /* [n_out][width] */
  int32_t ntaps, n_in, n_out, width;
  int32_t cap_q0, cap_w; /* entries available behind q0_out / w_out (>= n_out, >= n_out*width) */
} fsg_sepcompose_job;
int fsg_sep_compose(const fsg_sepcompose_job* jobs_host, int njobs, void* stream);

/* K4b — trilinear resampling onto a regular coarse grid + additive noise.
 * Replaces RandResample.__call__'s interpolation (synthseg.py:84-107 ->
 * utils/generation.py:227-285) and RandNoise (synthseg.py:217-235). */
typedef struct fsg_resample_job {
  const float* src;       /* [sx][sy][sz] */
  float* dst;             /* [n[0]][n[1]][n[2]] */
  const fsg_tab* tab[3];  /* lengths n[0], n[1], n[2] */
  const float* noise;     /* [n0*n1*n2] injected draws or NULL -> Philox (when has_noise) */
  fsg_rng rng;
  float noise_std;
  int32_t has_noise;
  int32_t n[3];           /* coarse grid of this job (jobs of one launch may differ) */
  int32_t _pad;
} fsg_resample_job;
int fsg_resample(const fsg_resample_job* jobs_host, int njobs, int sx, int sy, int sz, void* stream);

/* Elementwise x + std*N, clamp >= 0 at any resolution (RandNoise when no resampling ran). */
typedef struct fsg_noise_job {
  const float* src;
  float* dst;
  const float* noise;
  fsg_rng rng;
  float noise_std;
  int32_t flags; /* bit 0: do not clamp at 0 (multi-scale noise accumulation of StructNoise) */
} fsg_noise_job;
int fsg_add_noise(const fsg_noise_job* jobs_host, int njobs, int64_t nvox, void* stream);

/* K4c — separable linear zoom (myzoom_torch, utils/generation.py:310-397) with the global
 * reductions that follow it on this path.
 * post 0: raw zoom.  post 1: divide by the global max (RandResample.resize_back,
 * synthseg.py:109-114).  post 2: post 1 followed by ScaleIntensity(0,1)
 * (data/datasets.py:311).  For post>0 call fsg_zoom_minmax first; both read src only. */
typedef struct fsg_zoom_job {
  const float* src;       /* [n[0]][n[1]][n[2]] */
  float* dst;             /* [sx][sy][sz] */
  const fsg_tab* tab[3];  /* lengths sx, sy, sz */
  float* minmax;          /* [2] device: min, max of the zoomed volume */
  int32_t n[3];           /* source grid of this job (jobs of one launch may differ) */
  int32_t post;
} fsg_zoom_job;
int fsg_zoom_minmax(const fsg_zoom_job* jobs_host, int njobs, int sx, int sy, int sz, void* stream);
int fsg_zoom(const fsg_zoom_job* jobs_host, int njobs, int sx, int sy, int sz, void* stream);

/* ScaleIntensity(minv=0,maxv=1) standalone: reduce, then (x-min)/(max-min)
 * (data/datasets.py:40,311). minmax is a [2] device scratch. */
int fsg_minmax(const float* x, int64_t n, float* minmax, void* stream);
int fsg_scale_intensity(const float* x, float* out, int64_t n, const float* minmax, void* stream);

/* Label / dtype plumbing at the API edge (data/datasets.py:315-323). */
int fsg_f32_to_u8(const float* x, uint8_t* out, int64_t n, void* stream);
int fsg_u8_to_f32(const uint8_t* x, float* out, int64_t n, void* stream);
int fsg_u8_to_i64(const uint8_t* x, int64_t* out, int64_t n, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K5 — SR-artifact building blocks (generator/augmentation/artifacts.py, generator/artifacts/utils.py).
 * Single-volume calls (the artifacts run per sample after the batched base pipeline). */

/* Mixture of anisotropic Gaussians clamp(sum_k exp(-d_k^2/2), 0, 1) (mog_3d_tensor,
 * artifacts/utils.py:125-160).  centers/sigmas are [n][3] device arrays in AXIS order (axis 0,
 * 1, 2): the host applies the reference's x<->last-axis unpacking.  out (optional) receives the
 * weight g; if dst is given the BlurCortex blend dst = a*(1-g) + b*g is fused
 * (augmentation/artifacts.py:124-126). */
int fsg_mog(const float* centers, const float* sigmas, int n, int sx, int sy, int sz, float* out, const float* a, const float* b, float* dst, void* stream);

/* Weighted sampling WITHOUT replacement of k voxel centres among the voxels selected by a label
 * predicate (labels == match; match < 0: labels > 0; labels2 != NULL: labels2 > 0 && labels == 0),
 * weights = a small MoG prior (nprior <= 4, axis order) or uniform.  Same distribution as
 * torch.multinomial(p, k) / randperm(n)[:k] (artifacts.py:99-104, 558-560): exponential-race keys
 * -ln(u)/w from Philox, the k smallest win.  centers_out: [k][3] floats (axis order), count_out:
 * number actually drawn (min(k, candidates)).  workspace: >= 64 + 2048*8 bytes of device memory. */
typedef struct fsg_sample_job {
  const uint8_t* labels;
  const uint8_t* labels2;
  float* centers_out;
  int32_t* count_out;
  void* workspace;
  int64_t workspace_bytes;
  fsg_rng rng;
  float prior_centers[4][3];
  float prior_sigmas[4][3];
  int32_t match;
  int32_t nprior;
  int32_t k;
  int32_t _pad; /* != 0: write centres transposed (axis2, axis1, axis0), the order mog_3d_tensor's
                   (x0, y0, z0) unpacking gives to voxel indices (artifacts/utils.py:137-156) */
} fsg_sample_job;
int fsg_sample_voxels(const fsg_sample_job* job_host, int sx, int sy, int sz, void* stream);

/* Fractal Perlin noise sum_o amp_o * perlin_o (generate_fractal_noise_3d / generate_perlin_noise_3d,
 * artifacts/utils.py:224-388) + its global min/max (minmax: [2] device floats).  Per octave the
 * host supplies the unit gradient lattice [(r0+1)][(r1+1)][(r2+1)][3] (tile wrap applied) and the
 * per-axis lattice coordinates of every voxel (torch.linspace(0, res, S)). */
typedef struct fsg_perlin_octave {
  const float* grad;
  const float* lin[3];
  int32_t res[3];
  float amp;
} fsg_perlin_octave;
int fsg_perlin(const fsg_perlin_octave* octaves_host, int noct, int sx, int sy, int sz, float* out, float* minmax, void* stream);

/* StructNoise merge (artifacts.py:322-339).  scal: 6 device floats = min/max of the multi-scale
 * noise, min/max of the image, min/max of the raw fractal noise. */
int fsg_struct_blend(const float* x, const uint8_t* seg, const float* lr, const float* perlin, const float* scal, float noise_std, float increase, float* out, int64_t n,
                     void* stream);

/* Exact binary morphology on uint8 volumes (dilate/erode/apply_kernel, artifacts/utils.py:163-210;
 * build_halo / generate_fuzzy_boundaries / dilate_stack, artifacts.py:484-602).
 * fsg_morph_box: k^3 box, op 0 = dilation, 1 = erosion (zero padded), 2 = count (k^3 <= 255).
 * fsg_morph_dist: metric 0 = squared Euclidean distance to the mask (ball(r) dilation is dist <= r^2),
 *                 metric 1 = L1 distance (k 6-neighbour dilations are dist <= k); window r per axis.
 * fsg_morph_ring: (dilated && !mask), sub-sampled by an injected keep-mask or Bernoulli(p) (Philox).
 * fsg_morph_count_merge: out = mask || count > thr.
 * fsg_boundary_select: final mask and image of SimulatedBoundaries (artifacts.py:563-604). */
int fsg_morph_box(const uint8_t* src, uint8_t* dst, uint8_t* tmp, int k, int op, int sx, int sy, int sz, void* stream);
int fsg_morph_dist(const uint8_t* mask, uint16_t* dist_out, uint16_t* tmp, int r, int metric, int sx, int sy, int sz, void* stream);
int fsg_morph_thresh(const uint16_t* dist, uint8_t* out, int thr, int64_t n, void* stream);
int fsg_morph_ring(const uint8_t* dilated, const uint8_t* mask, const uint8_t* keep, fsg_rng rng, float p, uint8_t* out, int64_t n, void* stream);
int fsg_morph_count_merge(const uint8_t* count, const uint8_t* mask, int thr, uint8_t* out, int64_t n, void* stream);
int fsg_boundary_select(const float* x, const uint8_t* halo, const uint8_t* modif, const uint16_t* l1, const float* mog, int len, float* out, uint8_t* mask_out, int64_t n,
                        void* stream);
int fsg_mask_mul(const float* x, const uint8_t* mask, float* out, int64_t n, void* stream);
int fsg_label_mask(const uint8_t* labels, int match, uint8_t* out, int64_t n, void* stream);

/* K5-motion — SimulateMotion (artifacts.py:345-425): slice acquisition and PSF reconstruction.
 * Replace the reference's pybind11 module slice_acq_cuda (slice_acq_cuda.cpp:156-161) for the two
 * modes the generator uses:
 *   forward (slice_acq_cuda_kernel.cu:17-171) with interp_psf = false, no masks, no weights
 *     — slice_acq_cuda.forward(transforms, vol, [], [], psf, (h, w), res_slice, false, false);
 *   adjoint (:472-693) with interp_psf = true, equalize = true, no masks
 *     — slice_acq_cuda.adjoint_forward(transforms, psf, slices, [], [], (D,H,W), res_slice, true, true).
 * transforms: [n][3][4] row-major device floats.  vol: [D][H][W] (W fastest).  taps: [ntaps][4]
 * device floats = the NON-ZERO PSF entries in the reference's loop order as (ix_p, iy_p, iz_p, value),
 * 16-byte aligned; radius >= max |tap offset| (pixels farther than that from the volume are skipped).
 * slices: [n][h][w].  fsg_slice_acq_adjoint zero-fills its accumulator itself. */
int fsg_slice_acq_forward(const float* transforms, const float* vol, const float* taps, int ntaps, float radius, float* slices, int n, int h, int w, int D, int H, int W,
                          float res_slice, void* stream);
/* Acquisition from the x-pair volume: pairs[2 i] = vol[i], pairs[2 i + 1] = vol[i + 1] (built once per
 * volume by fsg_volume_xpairs; a Scanner acquires 2-6 stacks from the same volume).  Same result as
 * fsg_slice_acq_forward bit for bit; the two x corners of every trilinear sample arrive in one 8-byte
 * load, which halves the gather instructions of a kernel bound by L1 wavefronts. */
int fsg_volume_xpairs(const float* vol, float* pairs, int64_t nvox, void* stream);
/* ... and from the xy-quad volume (v[i], v[i+1], v[i+row_len], v[i+row_len+1]; row_len = W): two 16-byte loads per sample. */
int fsg_volume_xyquads(const float* vol, float* quads, int64_t nvox, int row_len, void* stream);
int fsg_slice_acq_forward_xyquads(const float* transforms, const float* vol_quads, const float* taps, int ntaps, float radius, float* slices, int n, int h, int w, int D, int H,
                                  int W, float res_slice, void* stream);
int fsg_slice_acq_forward_xpairs(const float* transforms, const float* vol_pairs, const float* taps, int ntaps, float radius, float* slices, int n, int h, int w, int D, int H,
                                 int W, float res_slice, void* stream);
int fsg_slice_acq_adjoint(const float* transforms, const float* psf, int dp, int hp, int wp, const float* taps, int ntaps, float radius, const float* slices,
                          const int32_t* slice_idx, float* vol, float* vol_weight, float* workspace, int n, int h, int w, int D, int H, int W, float res_slice, int equalize,
                          void* stream);
/* workspace: 2*D*H*W device floats, 8-byte aligned (interleaved value/weight accumulator: one 64-bit
 * reduction per tap); vol_weight may be NULL. */

/* The pybind module's FULL contract (any option the generator does not use): optional byte masks over the volume
 * ([D][H][W], non-zero = inside) and over the slices ([n][h][w]), the per-pixel weight output (need_weight), both PSF
 * modes in both directions, optional equalisation.
 *   fsg_slice_acq_forward_ex  = slice_acq_cuda.forward(transforms, vol, vol_mask, slices_mask, psf, (h, w), res_slice,
 *                               need_weight = (slices_weight != NULL), interp_psf)      (slice_acq_cuda.cpp:61-79)
 *   fsg_slice_acq_adjoint_ex  = slice_acq_cuda.adjoint_forward(transforms, psf, slices, slices_mask, vol_mask,
 *                               (D, H, W), res_slice, interp_psf, equalize)             (slice_acq_cuda.cpp:105-124)
 * psf: the dense [dp][hp][wp] grid (read when interp_psf != 0); taps / radius as above.  Outputs are zero-filled by
 * the call; masks and slices_weight may be NULL; vol_weight may be NULL unless equalize != 0. */
int fsg_slice_acq_forward_ex(const float* transforms, const float* vol, const uint8_t* vol_mask, const uint8_t* slices_mask, const float* psf, int dp, int hp, int wp,
                             const float* taps, int ntaps, float radius, float* slices, float* slices_weight, int n, int h, int w, int D, int H, int W, float res_slice,
                             int interp_psf, void* stream);
int fsg_slice_acq_adjoint_ex(const float* transforms, const float* psf, int dp, int hp, int wp, const float* taps, int ntaps, float radius, const float* slices,
                             const uint8_t* slices_mask, const uint8_t* vol_mask, float* vol, float* vol_weight, int n, int h, int w, int D, int H, int W, float res_slice,
                             int interp_psf, int equalize, void* stream);
/* Rigid-transform conversions of the reference's second pybind module (transform_convert_cuda.cpp:27-51,
 * kernels transform_convert_cuda_kernel.cu:14-65,190-264): (n, 6) axis-angle (radians) + translation <-> (n, 3, 4). */
int fsg_axisangle2mat(const float* axisangle, float* mat, int n, void* stream);
int fsg_mat2axisangle(const float* mat, float* axisangle, int n, void* stream);
/* slice_idx (device, [n], may be NULL): slice in of the launch reads slices[slice_idx[in]] — the
 * "stacks[kept_idx]" gather of PSFReconstructor (simulate_reco.py:766-767) without a copy.
 *
 * Slice-stack helpers of Scanner.scan (simulate_reco.py:300-466), all in place on [count] floats:
 *   fsg_slice_sums    per-slice sums (:408), sums[n] on the device
 *   fsg_slice_gamma   300 (s/300)^gamma, then / max (:225-236); workspace = 1 device float
 *   fsg_slice_rician  sqrt((s + n1 sigma)^2 + (n2 sigma)^2) where s > threshold (:248-257);
 *                     noise1/noise2 NULL: Philox draws (two per pixel), else injected arrays
 *   fsg_slice_void    slice idx[k] *= 1 - A exp(...) (:273-297); params[k] = (yc, xc, theta, a, A, sx) */
int fsg_slice_sums(const float* slices, int n, int hw, float* sums, void* stream);
int fsg_slice_gamma(float* slices, int64_t count, float gamma, float* workspace, void* stream);
int fsg_slice_rician(float* slices, int64_t count, float threshold, float sigma, const float* noise1, const float* noise2, fsg_rng rng, void* stream);
int fsg_slice_void(float* slices, int h, int w, const int32_t* idx, const float* params, int nvoid, void* stream);
/* Optional 3^3 mean (smooth_volume, simulate_reco.py:584-595) + merge with the clean volume
 * out = w*rec + (1-w)*gt (merge_volumes, :692-709); w = clamp((weight_raw + increase - min)/(max - min), 0, 1)
 * with minmax = [2] device floats (NULL: weight_raw is used as is); weight_raw NULL: no merge. */
int fsg_recon_merge(const float* rec, const float* gt, const float* weight_raw, const float* minmax, float increase, int smooth, int D, int H, int W, float* out, void* stream);

/* Copies nfloats (rounded up to 4) from pinned, device-mapped host memory into device memory with
 * an SM kernel instead of the copy engine: the per-step parameter block must not queue behind the
 * bulk H2D transfers of later pipeline steps.  Both pointers 16-byte aligned. */
int fsg_fetch_params(const float* src_host_mapped, float* dst, int64_t nfloats, void* stream);

/* Control grids of a batch drawn on the device: out[i] = scale * N(0,1) from the job's Philox
 * stream.  Used by the batched generator for the deformation control grid (Fsmall = nonlin_std *
 * randn, affine_nonrigid.py:312-316) and the bias control grid (synthseg.py:170-172) instead of a
 * host draw + upload per sample.  At most FSG_MAX_JOBS jobs per call. */
typedef struct fsg_grid_job {
  float* out;
  fsg_rng rng;
  float scale;
  int32_t n;
} fsg_grid_job;
int fsg_draw_grids(const fsg_grid_job* jobs_host, int njobs, void* stream);

/* Native launch builder of the batched base path.  One call builds every job struct of a step of B samples from
 * the per-sample draws and issues the step's launches — what FetalSynthGen.generate / augment (model.py:94-229)
 * drive for one sample, for a whole batch: fsg_gmm, fsg_draw_grids, fsg_warp_shift, fsg_warp, fsg_sep_compose,
 * fsg_sepconv, fsg_zoom_minmax, fsg_zoom (samples without the resolution simulation: fsg_add_noise and, with
 * `scale`, fsg_minmax + fsg_scale_intensity).  The host mirror (batch_step.py / engine.py) builds the same bytes in
 * Python; fsg_step_build exposes the job arrays so that tests can compare the two without a GPU.
 * Pointers marked HOST are read during the call; every other pointer is a device address. */
typedef struct fsg_step_sample {
  const uint8_t* seg;        /* [nvox] input segmentation */
  const void* words;         /* bit-packed seed words (see fsg_gmm_job) or NULL */
  const int8_t* seed[4];     /* label volumes when words == NULL (unused entries NULL) */
  int32_t word_bytes;
  int32_t shift[4];
  int32_t mask[4];
  int32_t deform, flip, gamma_on, bias_on, res_on, noise_on; /* per-sample gates */
  uint64_t sample_id;        /* Philox subsequence */
  const float* mus;          /* HOST [nlabels] */
  const float* sigmas;       /* HOST [nlabels] */
  float A[9];
  float c2[3];
  float nonlin_std, bf_std, gamma, noise_std;
  int32_t fs[3];             /* control grid of the deformation */
  int32_t bs[3];             /* control grid of the bias field */
  uint64_t tex, surf;        /* fsg_texvol handles of the sample's intensity volume, or 0: linear hand-over in buf[0] */
  const fsg_tab* ftab[3];    /* zoom tables control grid -> volume, per axis */
  const fsg_tab* btab[3];
  /* resolution simulation (res_on): */
  const fsg_tab* pos[3];     /* down-sampling positions per axis */
  const fsg_tab* ztab[3];    /* zoom tables coarse grid -> volume */
  int32_t n_out[3];          /* coarse extents */
  int32_t ntaps[3];          /* Gaussian taps per axis (1 = no blur) */
  const float* taps[3];      /* HOST tap arrays, NULL = no blur; equal pointers within a sample share one upload */
} fsg_step_sample;
typedef struct fsg_step {
  int32_t B, nlabels;
  int32_t shape[3];
  int32_t scale;             /* ScaleIntensity fused into the last kernel */
  float center[3];
  int32_t _pad;
  uint64_t seed;             /* Philox key */
  float* buf[3];             /* three scratch volumes per sample: rows at buf_pitch[k] bytes */
  int64_t buf_pitch[3];
  float* out_img;            /* [B][nvox] */
  uint8_t* out_seg;          /* [B][nvox] */
  float* grids;              /* control grids drawn on the device: rows of grids_cap floats */
  int64_t grids_pitch, grids_cap;
  float* shift;              /* [B] rows of >= 3 floats */
  int64_t shift_pitch;
  float* sep_tables;         /* composed axis tables: rows of sep_cap floats */
  int64_t sep_pitch, sep_cap;
  float* minmax;             /* [B] rows of >= 2 floats */
  int64_t minmax_pitch;
  float* ring_host;          /* HOST, pinned and device-mapped: parameter block of this step */
  float* ring_dev;
  int64_t ring_floats;
} fsg_step;
typedef struct fsg_step_jobs {
  int32_t n_gmm[2], n_grid, n_warp, n_shift, n_sep, n_noise, n_scale, ring_used, _pad;
  fsg_gmm_job gmm[2][FSG_MAX_JOBS];   /* [0]: packed seed words, [1]: label volumes */
  fsg_grid_job grid[2 * FSG_MAX_JOBS];
  fsg_warp_job warp[FSG_MAX_JOBS];
  fsg_warp_job shift[FSG_MAX_JOBS];
  fsg_sepcompose_job compose[3 * FSG_MAX_JOBS];
  fsg_sepconv_job sep[FSG_MAX_JOBS];
  fsg_zoom_job zoom[FSG_MAX_JOBS];
  fsg_noise_job noise[FSG_MAX_JOBS];
  int32_t scale_idx[FSG_MAX_JOBS];    /* samples that need the stand-alone ScaleIntensity */
} fsg_step_jobs;
/* Returns 0, or -1 when the step needs something this builder does not cover (the caller takes the generic
 * path): more than 31 Gaussian taps, a coarse grid larger than the volume, label-volume samples with different
 * numbers of volumes. */
int fsg_step_build(const fsg_step* step, const fsg_step_sample* samples_host, fsg_step_jobs* out);
int fsg_step_run(const fsg_step* step, const fsg_step_sample* samples_host, void* stream);

/* Per-sample parameter draws of the batched path, natively: every parameter FetalSynthGen.sample draws for a sample
 * (rand_gmm.py:81-85,120-145; affine_nonrigid.py:140-145,249-324; synthseg.py:63-80,157-176,217-235,262-275), for B
 * samples at once, as a pure function of (base_seed, sample id): uniform (sample, column) = splitmix64 counter
 * stream, shaped by the distributions of the stage objects.  The host mirror (batch_draw.draw_batch) calls this, so the
 * per-sample plans of the generic path and the native step share one definition.  All pointers are HOST memory. */
typedef struct fsg_draw_config {
  int32_t nlabels, nseed;              /* max(seed_labels) + 1, len(seed_labels) */
  const int32_t* seed_labels;          /* [nseed] */
  const int32_t* generation_classes;   /* [nseed] */
  int32_t tied;                        /* generation_classes != seed_labels: means tied to the class mean */
  int32_t meta_labels, min_subclusters, max_subclusters;
  int32_t shape[3];
  int32_t nonlinear;
  double res[3];
  double deform_prob, flip_prb, max_rotation, max_shear, max_scaling, nonlin_scale_min, nonlin_scale_max, nonlin_std_max;
  double centre2[3], max_shift[3];
  double gamma_prob, gamma_std;
  double bias_prob, bf_scale_min, bf_scale_max, bf_std_min, bf_std_max;
  double res_prob, min_resolution, max_resolution;
  double noise_prob, noise_std_min, noise_std_max;
} fsg_draw_config;
typedef struct fsg_draw_out {          /* struct of arrays over the batch, caller-allocated */
  float* mus;                          /* [B][nlabels] */
  float* sigmas;                       /* [B][nlabels] */
  uint8_t *deform_on, *flip, *gamma_on, *bias_on, *res_on, *noise_on; /* [B] */
  double *rot, *shear, *scal;          /* [B][3] */
  float* A;                            /* [B][9] */
  double* c2;                          /* [B][3] */
  double* nonlin_scale;                /* [B] */
  int64_t* size_f;                     /* [B][3] */
  float* nonlin_std;                   /* [B] */
  double* gamma;                       /* [B] */
  double* bf_scale;                    /* [B] */
  int64_t* bf_size;                    /* [B][3] */
  float* bf_std;                       /* [B] */
  double* spacing;                     /* [B] */
  double* stds;                        /* [B][3] */
  float* noise_std;                    /* [B] */
  int64_t* m2s;                        /* [B][meta_labels] or NULL */
} fsg_draw_out;
int fsg_draw_batch(const fsg_draw_config* cfg, const uint64_t* sample_ids, int B, uint64_t base_seed, fsg_draw_out* out);
/* Zero-padded Gaussian taps of make_gaussian_kernel (utils/generation.py:74-81) for one sigma: writes
 * 2 * ceil(3 sigma) + 1 normalised float32 taps into `out` (capacity `cap`), returns their number (or -1). */
int fsg_gaussian_taps(double sigma, float* out, int cap);

/* Inputs of a drawn step that are not drawn: where each sample's volumes live and the host mirror's caches of
 * device tables.  All pointers HOST arrays; entries are device addresses stored as 64-bit integers. */
typedef struct fsg_step_inputs {
  const uint64_t* seg;         /* [B] uint8 segmentation volumes */
  const uint64_t* words;       /* [B] bit-packed seed words of the sample's subject, 0 = label volumes instead */
  const int32_t* word_bytes;   /* [B] */
  const int32_t* const* layout; /* [B] -> the subject's field layout [nmax + 1][2] = (shift, mask) per sub-class count (-1: absent) */
  const int32_t* layout_len;   /* [B] nmax + 1 */
  const uint64_t* seed;        /* [B][4] label volumes when words[b] == 0 */
  const uint64_t* tex;         /* [B] fsg_texvol handles of the engine's block-linear volumes (0: linear hand-over) */
  const uint64_t* surf;        /* [B] */
  /* dense address tables per axis: [key] = device address of the cached 1-D table, 0 = not built yet */
  const uint64_t* zoom_tab[3]; /* key = control-grid extent -> zoom table to the volume extent */
  int32_t zoom_len[3];
  const uint64_t* pos_tab[3];  /* key = coarse extent -> down-sampling positions */
  const uint64_t* back_tab[3]; /* key = coarse extent -> zoom table back to the volume extent */
  int32_t res_len[3];
  float* taps_host;            /* [B][3][FSG_STEP_MAX_TAPS] scratch for the Gaussian taps of the step */
} fsg_step_inputs;
#define FSG_STEP_MAX_TAPS 64
/* Fills samples_out[B] from the draws + inputs (what batch_step.fill_step does in numpy).  Returns 0, -1 when the step
 * is not covered by the native builder, or -2 when a table is missing: then missing_out[0..2] = (kind 0 zoom / 1 pos+back,
 * axis, key) of the first missing table, for the caller to build and retry. */
int fsg_step_fill(const fsg_step* step, const fsg_draw_config* cfg, const fsg_draw_out* draw, const uint64_t* sample_ids, const fsg_step_inputs* in,
                  fsg_step_sample* samples_out, int32_t* missing_out);

/* Bit-packed seed cache.  A subject's seed volumes for every sub-class count
 * share the meta-label support, so one word per voxel holds them all: bits 0-2 the meta-label
 * (0 = background, 1..4), then one field per sub-class count n >= 2 holding the voxel's sub-class
 * index (ceil(log2 n) bits; 14 bits in total for counts 1..6 -> uint16, 28 bits for 1..10 -> uint32).
 * fsg_unpack_seeds writes the label volume a sample needs — what the sum of the four selected seed
 * files gives (rand_gmm.py:90-97): out[v] = meta ? 10 * meta + ((word >> shift[meta-1]) & mask[meta-1]) : 0
 * with shift / mask the field of the count drawn for that meta-label (mask 0 for one sub-class). */
typedef struct fsg_unpack_job {
  const void* words; /* [nvox] uint16 or uint32 (device) */
  uint8_t* out;      /* [nvox] (device) */
  int32_t shift[4];
  int32_t mask[4];
  int32_t word_bytes; /* 2 or 4 */
  int32_t _pad;
} fsg_unpack_job;
int fsg_unpack_seeds(const fsg_unpack_job* jobs_host, int njobs, int64_t nvox, void* stream);

/* ------------------------------------------------------------------------------------------
 * K6 — seed generation.  Replaces scripts/generate_seeds.py:133-211:
 * label fusion into meta-labels + sklearn GaussianMixture(n_components, n_init=5,
 * init_params="k-means++").fit_predict on the image voxels of each meta-label (one feature,
 * full covariance, float64 like the reference's call: sklearn up-casts the torch tensor it is given).
 *
 * fsg_seed_partition (generate_seeds.py:133-146,190-198): meta = lut[seg]; lut value 4 marks the
 *   labels treated as background, which become meta-label 4 where image != 0 and 0 elsewhere; NaN
 *   image values count as 0.  Writes the image values of meta-labels 1..4 back to back, each in
 *   voxel order, into x[n] with their flat voxel indices in index[n], and the four partition
 *   sizes into counts[4] (device).  lut: 256 bytes of HOST memory.  n < 2^31.
 *   workspace: >= fsg_seed_partition_workspace(n) bytes of device memory.
 * fsg_em_seed: greedy k-means++ (sklearn/cluster/_kmeans.py:_kmeans_plusplus, 2 + int(ln k) local
 *   trials, D^2 sampling by an exponential race over Philox uniforms) -> job.seeds[k] sample indices.
 * fsg_em_fit: _initialize from the one-hot responsibilities at job.seeds, then E / M steps until
 *   |delta lower bound| < tol or max_iter; every job is one initialisation, all jobs advance in the
 *   same launches and a converged job's blocks exit at once (no host round trip inside the loop).
 *   Results: job.params = weights | means | covariances (3 x FSG_EM_MAXK doubles), job.trace[i] =
 *   lower bound of iteration i (1-based; the caller keeps the initialisation with the largest final
 *   value, mixture/_base.py:282-287), job.state = {n_iter, converged}.
 * fsg_em_predict: final E step, argmax component (first maximum) -> job.labels[i] (if non-NULL)
 *   and job.out[job.index[i]] = label_base + component (if out non-NULL).
 * Any njobs >= 1; job structs are HOST memory; workspace: >= fsg_em_workspace(njobs) bytes. */
#define FSG_EM_MAXK 16
typedef struct fsg_em_job {
  const float* x;       /* [n] values of one meta-label (device) */
  const int32_t* index; /* [n] flat voxel index of every value (device); may be NULL when out is NULL */
  int64_t n;
  int32_t k;            /* components, 1..FSG_EM_MAXK, k <= n */
  int32_t label_base;   /* 10 * meta-label (generate_seeds.py:207-209) */
  double* params;       /* [3 * FSG_EM_MAXK] (device) */
  double* trace;        /* [max_iter + 1] (device) */
  int32_t* state;       /* [2] (device) */
  int32_t* seeds;       /* [FSG_EM_MAXK] initial sample indices (device): out of fsg_em_seed, in of fsg_em_fit */
  uint8_t* labels;      /* [n] or NULL (device) */
  int8_t* out;          /* seed volume or NULL (device) */
  uint64_t rng_seed;    /* Philox key of the k-means++ draws */
  uint64_t rng_stream;  /* ... and subsequence: one per initialisation */
} fsg_em_job;
int64_t fsg_seed_partition_workspace(int64_t n);
int fsg_seed_partition(const float* image, const uint8_t* seg, const uint8_t* lut_host, int64_t n, float* x, int32_t* index, int64_t* counts, void* workspace,
                       int64_t workspace_bytes, void* stream);
int64_t fsg_em_workspace(int njobs);
int fsg_em_seed(const fsg_em_job* jobs_host, int njobs, void* workspace, int64_t workspace_bytes, void* stream);
int fsg_em_fit(const fsg_em_job* jobs_host, int njobs, int max_iter, double tol, double reg_covar, void* workspace, int64_t workspace_bytes, void* stream);
int fsg_em_predict(const fsg_em_job* jobs_host, int njobs, void* workspace, int64_t workspace_bytes, void* stream);

/* RNG self-test: fills out[n] with Philox standard normals exactly as the kernels draw them;
 * raw != 0 writes the raw 32-bit words instead (for the Random123 known-answer test). */
int fsg_philox_fill(fsg_rng rng, float* out, int64_t n, int raw, void* stream);

int fsg_version(void);
const char* fsg_last_error(void);
int fsg_sizeof(const char* struct_name); /* ABI check for bindings */

#ifdef __cplusplus
}
#endif
#endif /* FSG_H */
