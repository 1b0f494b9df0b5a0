// Native launch builder of the batched base path (fsg_step_build / fsg_step_run, see fsg.h): host code only.
// It fills, for a whole batch, the job structs that engine.py / batch_step.py fill in Python — same fields, same
// buffer roles, same launch order — and issues the step's launches through the library's own entry points, so a
// step costs one C-ABI call.  Buffer roles of a sample b (rows of the three scratch volumes):
//   buf[0]: GMM image (linear hand-over only), later the y-pass scratch of the resolution simulation
//   buf[1]: deformed image (input of the resolution simulation / of the stand-alone noise)
//   buf[2]: coarse image (x-pass scratch and output of the resolution simulation, input of the zoom back)
#include <algorithm>
#include <cstring>

#include "common.cuh"

using namespace fsg;

namespace {

// Philox stream ids of the stages (engine.py: STAGE_GMM, STAGE_NOISE, STAGE_FIELD, STAGE_BIAS)
constexpr uint32_t kStageGmm = 1, kStageNoise = 2, kStageField = 3, kStageBias = 4;

inline int64_t ceil4(int64_t v) { return (v + 3) / 4 * 4; }
template <typename T>
inline T* row(T* base, int64_t pitch_bytes, int b) {
  return reinterpret_cast<T*>(reinterpret_cast<char*>(base) + pitch_bytes * b);
}
inline fsg_rng rng_of(uint64_t seed, uint64_t sample, uint32_t stage) {
  fsg_rng r;
  r.seed = seed;
  r.sample = sample;
  r.stage = stage;
  r._pad = 0;
  return r;
}

}  // namespace

extern "C" int fsg_step_build(const fsg_step* st, const fsg_step_sample* S, fsg_step_jobs* J) {
  FSG_REQUIRE(st && S && J, "fsg_step_build: NULL argument");
  const int B = st->B;
  FSG_REQUIRE(B >= 1 && B <= FSG_MAX_JOBS, "fsg_step_build: B=%d outside [1,%d]", B, FSG_MAX_JOBS);
  const int sx = st->shape[0], sy = st->shape[1], sz = st->shape[2];
  FSG_REQUIRE(sx >= 1 && sy >= 1 && sz >= 1, "fsg_step_build: bad shape");
  const int64_t nvox = (int64_t)sx * sy * sz;
  FSG_REQUIRE(st->buf[0] && st->buf[1] && st->buf[2] && st->out_img && st->out_seg && st->shift && st->minmax && st->ring_host && st->ring_dev,
              "fsg_step_build: NULL engine buffer");
  FSG_REQUIRE(st->nlabels >= 1 && st->nlabels <= 256, "fsg_step_build: nlabels=%d outside [1,256]", st->nlabels);
  memset(J, 0, sizeof(*J));

  // ---- parameter block: means [B][nlp], sigmas [B][nlp], then the Gaussian taps
  const int nl = st->nlabels;
  const int64_t nlp = ceil4(nl);
  int64_t used = 2 * B * nlp;
  FSG_REQUIRE(used <= st->ring_floats, "fsg_step_build: parameter ring too small");
  for (int b = 0; b < B; ++b) {
    FSG_REQUIRE(S[b].mus && S[b].sigmas && S[b].seg, "fsg_step_build: sample %d has a NULL mus/sigmas/seg", b);
    memcpy(st->ring_host + b * nlp, S[b].mus, sizeof(float) * nl);
    memcpy(st->ring_host + (B + b) * nlp, S[b].sigmas, sizeof(float) * nl);
  }

  // ---- K1
  int nvol = -1;
  for (int b = 0; b < B; ++b) {
    const fsg_step_sample& s = S[b];
    const int kind = s.words ? 0 : 1;
    fsg_gmm_job& g = J->gmm[kind][J->n_gmm[kind]++];
    g.mus = st->ring_dev + b * nlp;
    g.sigmas = st->ring_dev + (B + b) * nlp;
    g.nlabels = nl;
    g.rng = rng_of(st->seed, s.sample_id, kStageGmm);
    if (s.words) {
      g.words = s.words;
      g.word_bytes = s.word_bytes;
      for (int m = 0; m < 4; ++m) g.shift[m] = s.shift[m], g.mask[m] = s.mask[m];
    } else {
      int n = 0;
      for (int m = 0; m < 4; ++m) {
        g.seed[m] = s.seed[m];
        n += s.seed[m] != nullptr;
      }
      if (n < 1 || (nvol >= 0 && n != nvol)) return -1;  // a launch needs one number of label volumes
      nvol = n;
    }
    if (s.surf) {
      g.out_surf = s.surf;
      g.row_len = sz;
      g.surf_ny = sy;
    } else {
      g.out = row(st->buf[0], st->buf_pitch[0], b);
    }
  }

  // ---- control grids drawn on the device: per sample the field grid, then the bias grid
  int64_t nf = 0, nb = 0;
  for (int b = 0; b < B; ++b) {
    if (S[b].deform) nf = std::max<int64_t>(nf, 3LL * S[b].fs[0] * S[b].fs[1] * S[b].fs[2]);
    if (S[b].bias_on) nb = std::max<int64_t>(nb, (int64_t)S[b].bs[0] * S[b].bs[1] * S[b].bs[2]);
  }
  nf = ceil4(nf);
  nb = ceil4(nb);
  if (nf + nb > 0) FSG_REQUIRE(st->grids && nf + nb <= st->grids_cap, "fsg_step_build: control-grid buffer too small (%lld floats needed)", (long long)(nf + nb));
  for (int b = 0; b < B; ++b) {
    const fsg_step_sample& s = S[b];
    if (s.deform) {
      FSG_REQUIRE(s.fs[0] >= 1 && s.fs[1] >= 1 && s.fs[2] >= 1, "fsg_step_build: sample %d: bad control grid", b);
      fsg_grid_job& q = J->grid[J->n_grid++];
      q.out = row(st->grids, st->grids_pitch, b);
      q.n = 3 * s.fs[0] * s.fs[1] * s.fs[2];
      q.scale = s.nonlin_std;
      q.rng = rng_of(st->seed, s.sample_id, kStageField);
    }
    if (s.bias_on) {
      FSG_REQUIRE(s.bs[0] >= 1 && s.bs[1] >= 1 && s.bs[2] >= 1, "fsg_step_build: sample %d: bad bias grid", b);
      fsg_grid_job& q = J->grid[J->n_grid++];
      q.out = row(st->grids, st->grids_pitch, b) + nf;
      q.n = s.bs[0] * s.bs[1] * s.bs[2];
      q.scale = s.bf_std;
      q.rng = rng_of(st->seed, s.sample_id, kStageBias);
    }
  }

  // ---- K2
  J->n_warp = B;
  for (int b = 0; b < B; ++b) {
    const fsg_step_sample& s = S[b];
    fsg_warp_job& w = J->warp[b];
    if (s.tex)
      w.src_tex = s.tex;
    else
      w.src_img = row(st->buf[0], st->buf_pitch[0], b);
    w.src_seg = s.seg;
    w.dst_seg = st->out_seg + (int64_t)b * nvox;
    const bool direct = !s.res_on && !s.noise_on;  // nothing follows the deformation: write the output at once
    w.dst_img = direct ? st->out_img + (int64_t)b * nvox : row(st->buf[1], st->buf_pitch[1], b);
    w.mode = s.deform ? 1 : 0;
    w.flip = (s.deform && s.flip) ? 1 : 0;
    w.shift = row(st->shift, st->shift_pitch, b);
    if (s.deform) {
      for (int q = 0; q < 9; ++q) w.A[q] = s.A[q];
      for (int q = 0; q < 3; ++q) w.c2[q] = s.c2[q], w.center[q] = st->center[q], w.fs[q] = s.fs[q], w.ftab[q] = s.ftab[q];
      w.fsmall = row(st->grids, st->grids_pitch, b);
    }
    if (s.gamma_on) {
      w.has_gamma = 1;
      w.gamma = s.gamma;
    }
    if (s.bias_on) {
      w.bf_low = row(st->grids, st->grids_pitch, b) + nf;
      for (int q = 0; q < 3; ++q) w.bs[q] = s.bs[q], w.btab[q] = s.btab[q];
    }
    if (s.deform) J->shift[J->n_shift++] = w;
  }

  // ---- K4: resolution simulation and the zoom back
  int R = 0;
  int64_t maxw = 2, nmax = std::max(sx, std::max(sy, sz));
  for (int b = 0; b < B; ++b) {
    if (!S[b].res_on) continue;
    ++R;
    for (int a = 0; a < 3; ++a) {
      FSG_REQUIRE(S[b].n_out[a] >= 1 && S[b].pos[a] && S[b].ztab[a] && S[b].ntaps[a] >= 1, "fsg_step_build: sample %d: incomplete resolution tables", b);
      if (S[b].n_out[a] > st->shape[a]) return -1;  // an up-sampled axis needs larger scratch rows: generic path
      nmax = std::max<int64_t>(nmax, S[b].n_out[a]);
      if (S[b].taps[a]) maxw = std::max<int64_t>(maxw, S[b].ntaps[a] + 1);
    }
  }
  if (R) {
    maxw = std::max<int64_t>(32, ceil4(maxw));
    const int64_t per_axis = ceil4(nmax * maxw + (nmax + 1) / 2);
    FSG_REQUIRE(st->sep_tables && 3 * per_axis <= st->sep_cap, "fsg_step_build: axis-table buffer too small (%lld floats needed)", (long long)(3 * per_axis));
    int k = 0;
    for (int b = 0; b < B; ++b) {
      const fsg_step_sample& s = S[b];
      if (!s.res_on) continue;
      // taps of this sample: one upload per distinct array
      const float* tdev[3] = {nullptr, nullptr, nullptr};
      for (int a = 0; a < 3; ++a) {
        if (!s.taps[a]) continue;
        for (int p = 0; p < a; ++p)
          if (s.taps[p] == s.taps[a]) tdev[a] = tdev[p];
        if (!tdev[a]) {
          FSG_REQUIRE(used + ceil4(s.ntaps[a]) <= st->ring_floats, "fsg_step_build: parameter ring too small");
          memcpy(st->ring_host + used, s.taps[a], sizeof(float) * s.ntaps[a]);
          tdev[a] = st->ring_dev + used;
          used += ceil4(s.ntaps[a]);
        }
      }
      fsg_sepconv_job& j = J->sep[k];
      float* ws = row(st->sep_tables, st->sep_pitch, k);
      for (int a = 0; a < 3; ++a) {
        fsg_sepcompose_job& c = J->compose[3 * k + a];
        const int n_in = st->shape[a];
        const int ntaps = s.taps[a] ? s.ntaps[a] : 1;
        const int width = std::min(n_in, ntaps + 1);
        float* w_ptr = ws + a * per_axis;
        int16_t* q_ptr = reinterpret_cast<int16_t*>(w_ptr + nmax * maxw);
        c.pos = s.pos[a];
        c.taps = tdev[a];
        c.q0_out = q_ptr;
        c.w_out = w_ptr;
        c.ntaps = ntaps;
        c.n_in = n_in;
        c.n_out = s.n_out[a];
        c.width = width;
        c.cap_q0 = (int32_t)nmax;
        c.cap_w = (int32_t)(nmax * maxw);
        fsg_sepaxis& ax = j.ax[a];
        ax.q0 = q_ptr;
        ax.w = w_ptr;
        ax.n_out = s.n_out[a];
        ax.width = width;
        ax.pos = s.pos[a];
        ax.taps = tdev[a];
        ax.ntaps = ntaps;
      }
      j.src = row(st->buf[1], st->buf_pitch[1], b);
      j.dst = row(st->buf[2], st->buf_pitch[2], b);
      j.tmp1 = j.dst;
      j.tmp2 = row(st->buf[0], st->buf_pitch[0], b);
      j.cap_dst = j.cap_tmp1 = j.cap_tmp2 = nvox;
      if (s.noise_on) {
        j.has_noise = 1;
        j.noise_std = s.noise_std;
        j.rng = rng_of(st->seed, s.sample_id, kStageNoise);
      }
      fsg_zoom_job& z = J->zoom[k];
      z.src = j.dst;
      z.dst = st->out_img + (int64_t)b * nvox;
      for (int a = 0; a < 3; ++a) z.tab[a] = s.ztab[a], z.n[a] = s.n_out[a];
      z.minmax = row(st->minmax, st->minmax_pitch, k);
      z.post = st->scale ? 2 : 1;
      ++k;
    }
  }
  J->n_sep = R;

  // ---- samples without the resolution simulation
  for (int b = 0; b < B; ++b) {
    const fsg_step_sample& s = S[b];
    if (s.res_on) continue;
    if (s.noise_on) {
      fsg_noise_job& n = J->noise[J->n_noise++];
      n.src = row(st->buf[1], st->buf_pitch[1], b);
      n.dst = st->out_img + (int64_t)b * nvox;
      n.noise_std = s.noise_std;
      n.rng = rng_of(st->seed, s.sample_id, kStageNoise);
    }
    if (st->scale) J->scale_idx[J->n_scale++] = b;
  }
  J->ring_used = (int32_t)used;
  return 0;
}

extern "C" int fsg_step_run(const fsg_step* st, const fsg_step_sample* S, void* stream) {
  static thread_local fsg_step_jobs J;
  if (int rc = fsg_step_build(st, S, &J)) return rc;
  const int sx = st->shape[0], sy = st->shape[1], sz = st->shape[2];
  const int64_t nvox = (int64_t)sx * sy * sz;
  int rc = 0;
  if ((rc = fsg_fetch_params(st->ring_host, st->ring_dev, J.ring_used, stream))) return rc;
  for (int k = 0; k < 2; ++k)
    if (J.n_gmm[k] && (rc = fsg_gmm(J.gmm[k], J.n_gmm[k], nvox, stream))) return rc;
  for (int q = 0; q < J.n_grid; q += FSG_MAX_JOBS)
    if ((rc = fsg_draw_grids(J.grid + q, std::min(FSG_MAX_JOBS, J.n_grid - q), stream))) return rc;
  if (J.n_shift && (rc = fsg_warp_shift(J.shift, J.n_shift, sx, sy, sz, stream))) return rc;
  if ((rc = fsg_warp(J.warp, J.n_warp, sx, sy, sz, stream))) return rc;
  if (J.n_sep) {
    if ((rc = fsg_sep_compose(J.compose, 3 * J.n_sep, stream))) return rc;
    if ((rc = fsg_sepconv(J.sep, J.n_sep, sx, sy, sz, stream))) return rc;
    if ((rc = fsg_zoom_minmax(J.zoom, J.n_sep, sx, sy, sz, stream))) return rc;
    if ((rc = fsg_zoom(J.zoom, J.n_sep, sx, sy, sz, stream))) return rc;
  }
  if (J.n_noise && (rc = fsg_add_noise(J.noise, J.n_noise, nvox, stream))) return rc;
  for (int i = 0; i < J.n_scale; ++i) {
    float* x = st->out_img + (int64_t)J.scale_idx[i] * nvox;
    float* mm = reinterpret_cast<float*>(reinterpret_cast<char*>(st->minmax) + st->minmax_pitch * (J.n_sep + i));
    if ((rc = fsg_minmax(x, nvox, mm, stream))) return rc;
    if ((rc = fsg_scale_intensity(x, x, nvox, mm, stream))) return rc;
  }
  return 0;
}
