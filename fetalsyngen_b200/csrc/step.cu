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

// ------------------------------------------------------------------------------------------------ draws
namespace {

inline uint64_t splitmix(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
// uniform in (0, 1) of (base seed, sample id, column): 53 bits, never 0 or 1
struct Uniforms {
  uint64_t key;
  Uniforms(uint64_t base_seed, uint64_t id) : key(splitmix(base_seed ^ splitmix(id))) {}
  double operator()(uint64_t col) const { return ((double)(splitmix(key + col * 0xD1342543DE82EF95ull) >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
};

// Inverse of the standard normal distribution function: Wichura's algorithm AS 241 (PPND16), relative error ~1e-16.
double ndtri(double p) {
  const double q = p - 0.5;
  if (std::fabs(q) <= 0.425) {
    const double r = 0.180625 - q * q;
    const double num = (((((((2.5090809287301226727e3 * r + 3.3430575583588128105e4) * r + 6.7265770927008700853e4) * r + 4.5921953931549871457e4) * r + 1.3731693765509461125e4) * r +
                          1.9715909503065514427e3) * r + 1.3314166789178437745e2) * r + 3.3871328727963666080e0);
    const double den = (((((((5.2264952788528545610e3 * r + 2.8729085735721942674e4) * r + 3.9307895800092710610e4) * r + 2.1213794301586595867e4) * r + 5.3941960214247511077e3) * r +
                          6.8718700749205790830e2) * r + 4.2313330701600911252e1) * r + 1.0);
    return q * num / den;
  }
  double r = q < 0 ? p : 1.0 - p;
  r = std::sqrt(-std::log(r));
  double v;
  if (r <= 5.0) {
    r -= 1.6;
    const double num = (((((((7.74545014278341407640e-4 * r + 2.27238449892691845833e-2) * r + 2.41780725177450611770e-1) * r + 1.27045825245236838258e0) * r + 3.64784832476320460504e0) * r +
                          5.76949722146069140550e0) * r + 4.63033784615654529590e0) * r + 1.42343711074968357734e0);
    const double den = (((((((1.05075007164441684324e-9 * r + 5.47593808499534494600e-4) * r + 1.51986665636164571966e-2) * r + 1.48103976427480074590e-1) * r + 6.89767334985100004550e-1) * r +
                          1.67638483018380384940e0) * r + 2.05319162663775882187e0) * r + 1.0);
    v = num / den;
  } else {
    r -= 5.0;
    const double num = (((((((2.01033439929228813265e-7 * r + 2.71155556874348757815e-5) * r + 1.24266094738807843860e-3) * r + 2.65321895265761230930e-2) * r + 2.96560571828504891230e-1) * r +
                          1.78482653991729133580e0) * r + 5.46378491116411436990e0) * r + 6.65790464350110377720e0);
    const double den = (((((((2.04426310338993978564e-15 * r + 1.42151175831644588870e-7) * r + 1.84631831751005468180e-5) * r + 7.86869131145613259100e-4) * r + 1.48753612908506148525e-2) * r +
                          1.36929880922735805310e-1) * r + 5.99832206555887937690e-1) * r + 1.0);
    v = num / den;
  }
  return q < 0 ? -v : v;
}

inline void matmul3(const double* a, const double* b, double* c) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) c[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
}

}  // namespace

extern "C" int fsg_draw_batch(const fsg_draw_config* c, const uint64_t* ids, int B, uint64_t base_seed, fsg_draw_out* o) {
  FSG_REQUIRE(c && ids && o && B >= 1, "fsg_draw_batch: bad arguments");
  FSG_REQUIRE(c->nlabels >= 1 && c->nseed >= 1 && c->seed_labels && c->generation_classes, "fsg_draw_batch: bad label configuration");
  const int nl = c->nlabels, ns = c->nseed;
  const uint64_t o_s = 2ull * nl + ns;  // first column of the scalar block
  const double pi = 3.141592653589793;
  for (int b = 0; b < B; ++b) {
    const Uniforms U(base_seed, ids[b]);
    auto S = [&](int k) { return U(o_s + (uint64_t)k); };
    // ---- GMM tables (rand_gmm.py:120-145)
    float* mus = o->mus + (size_t)b * nl;
    float* sig = o->sigmas + (size_t)b * nl;
    for (int l = 0; l < nl; ++l) {
      mus[l] = 25.0f + 200.0f * (float)U((uint64_t)l);
      sig[l] = 5.0f + 20.0f * (float)U((uint64_t)(nl + l));
    }
    if (c->tied) {
      // every seed label's mean is its class mean plus N(0, 25), clipped; class means are read before any is replaced
      float t[256];
      FSG_REQUIRE(ns <= 256, "fsg_draw_batch: more than 256 seed labels");
      for (int k = 0; k < ns; ++k) t[k] = mus[c->generation_classes[k]] + 25.0f * (float)ndtri(U(2ull * nl + (uint64_t)k));
      for (int k = 0; k < ns; ++k) mus[c->seed_labels[k]] = std::min(std::max(t[k], 0.0f), 225.0f);
    }
    // ---- spatial deformation (affine_nonrigid.py:140-145, 249-324)
    o->deform_on[b] = S(0) < c->deform_prob;
    o->flip[b] = S(1) < c->flip_prb;
    double rot[3], sh[3], sc[3];
    for (int a = 0; a < 3; ++a) {
      rot[a] = (2 * c->max_rotation * S(2 + a) - c->max_rotation) / 180.0 * pi;
      sh[a] = 2 * c->max_shear * S(5 + a) - c->max_shear;
      sc[a] = 1 + (2 * c->max_scaling * S(8 + a) - c->max_scaling);
      o->rot[3 * b + a] = rot[a], o->shear[3 * b + a] = sh[a], o->scal[3 * b + a] = sc[a];
    }
    {
      // make_affine_matrix (utils/generation.py:39-71): shear_x . shear_y . shear_z . Rx . Ry . Rz, rows scaled
      const double c0 = std::cos(rot[0]), s0 = std::sin(rot[0]), c1 = std::cos(rot[1]), s1 = std::sin(rot[1]), c2 = std::cos(rot[2]), s2 = std::sin(rot[2]);
      const double m[6][9] = {{1, 0, 0, sh[1], 1, 0, sh[2], 0, 1}, {1, sh[0], 0, 0, 1, 0, 0, sh[2], 1}, {1, 0, sh[0], 0, 1, sh[1], 0, 0, 1},
                              {1, 0, 0, 0, c0, -s0, 0, s0, c0},    {c1, 0, s1, 0, 1, 0, -s1, 0, c1},    {c2, -s2, 0, s2, c2, 0, 0, 0, 1}};
      double acc[9], tmp[9];
      memcpy(acc, m[0], sizeof(acc));
      for (int k = 1; k < 6; ++k) {
        matmul3(acc, m[k], tmp);
        memcpy(acc, tmp, sizeof(acc));
      }
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) o->A[9 * b + 3 * i + j] = (float)(acc[3 * i + j] * sc[i]);
    }
    for (int a = 0; a < 3; ++a) o->c2[3 * b + a] = c->centre2[a] + (2 * (c->max_shift[a] * S(11 + a)) - c->max_shift[a]);
    o->nonlin_scale[b] = c->nonlin_scale_min + S(14) * (c->nonlin_scale_max - c->nonlin_scale_min);
    for (int a = 0; a < 3; ++a) o->size_f[3 * b + a] = (int64_t)std::nearbyint(o->nonlin_scale[b] * (double)c->shape[a]);
    o->nonlin_std[b] = (float)(c->nonlin_std_max * S(15));
    // ---- gamma (synthseg.py:262-275), bias field (:157-176), resolution (:63-80), noise (:217-235)
    o->gamma_on[b] = S(16) < c->gamma_prob;
    o->gamma[b] = std::exp(c->gamma_std * ndtri(S(17)));
    o->bias_on[b] = S(18) < c->bias_prob;
    o->bf_scale[b] = c->bf_scale_min + S(19) * (c->bf_scale_max - c->bf_scale_min);
    for (int a = 0; a < 3; ++a) o->bf_size[3 * b + a] = std::max<int64_t>((int64_t)std::nearbyint(o->bf_scale[b] * (double)c->shape[a]), 1);
    o->bf_std[b] = (float)(c->bf_std_min + (c->bf_std_max - c->bf_std_min) * S(20));
    o->res_on[b] = S(21) < c->res_prob;
    o->spacing[b] = c->min_resolution + (c->max_resolution - c->min_resolution) * S(22);
    for (int a = 0; a < 3; ++a) {
      const double sd = (0.85 + 0.3 * S(23)) * std::log(5.0) / pi * o->spacing[b] / c->res[a];
      o->stds[3 * b + a] = o->spacing[b] <= c->res[a] ? 0.0 : sd;
    }
    o->noise_on[b] = S(24) < c->noise_prob;
    o->noise_std[b] = (float)(c->noise_std_min + (c->noise_std_max - c->noise_std_min) * S(25));
    if (o->m2s) {  // rand_gmm.py:81-85: randint(min, max + 1) per meta label
      const int k = c->max_subclusters - c->min_subclusters + 1;
      for (int m = 0; m < c->meta_labels; ++m) o->m2s[(size_t)b * c->meta_labels + m] = c->min_subclusters + std::min((int)(S(26 + m) * k), k - 1);
    }
  }
  return 0;
}

extern "C" int fsg_gaussian_taps(double sigma, float* out, int cap) {
  FSG_REQUIRE(out && sigma > 0, "fsg_gaussian_taps: bad arguments");
  const int sl = (int)std::ceil(3 * sigma);
  const int n = 2 * sl + 1;
  if (n > cap) return -1;
  // float32 arithmetic like torch.linspace / exp in make_gaussian_kernel; the sum in index order
  const float s = (float)sigma;
  float sum = 0.f;
  for (int i = 0; i < n; ++i) {
    const float t = (float)(i - sl) / s;
    out[i] = std::exp(-(t * t) / 2.0f);
    sum += out[i];
  }
  for (int i = 0; i < n; ++i) out[i] /= sum;
  return n;
}

extern "C" int fsg_step_fill(const fsg_step* st, const fsg_draw_config* c, const fsg_draw_out* d, const uint64_t* ids, const fsg_step_inputs* in, fsg_step_sample* S,
                             int32_t* missing) {
  FSG_REQUIRE(st && c && d && ids && in && S && missing, "fsg_step_fill: NULL argument");
  const int B = st->B;
  FSG_REQUIRE(B >= 1 && B <= FSG_MAX_JOBS, "fsg_step_fill: B=%d outside [1,%d]", B, FSG_MAX_JOBS);
  const int nl = c->nlabels;
  memset(S, 0, sizeof(fsg_step_sample) * B);
  for (int b = 0; b < B; ++b) {
    fsg_step_sample& s = S[b];
    s.seg = reinterpret_cast<const uint8_t*>(in->seg[b]);
    if (in->words && in->words[b]) {
      s.words = reinterpret_cast<const void*>(in->words[b]);
      s.word_bytes = in->word_bytes[b];
      FSG_REQUIRE(d->m2s && c->meta_labels == 4, "fsg_step_fill: packed seeds need the sub-class counts of four meta-labels");
      for (int m = 0; m < 4; ++m) {
        const int64_t n = d->m2s[(size_t)b * 4 + m];
        FSG_REQUIRE(n >= 0 && n < in->layout_len[b] && in->layout[b][2 * n] >= 0, "fsg_step_fill: sample %d: no seeds with %lld sub-classes in the subject's cache", b, (long long)n);
        s.shift[m] = in->layout[b][2 * n];
        s.mask[m] = in->layout[b][2 * n + 1];
      }
    } else {
      for (int m = 0; m < 4; ++m) s.seed[m] = reinterpret_cast<const int8_t*>(in->seed[4 * b + m]);
    }
    s.deform = d->deform_on[b], s.flip = d->flip[b], s.gamma_on = d->gamma_on[b], s.bias_on = d->bias_on[b], s.res_on = d->res_on[b], s.noise_on = d->noise_on[b];
    if (s.deform && !c->nonlinear) return -1;
    s.sample_id = ids[b];
    s.mus = d->mus + (size_t)b * nl;
    s.sigmas = d->sigmas + (size_t)b * nl;
    for (int q = 0; q < 9; ++q) s.A[q] = d->A[9 * b + q];
    for (int q = 0; q < 3; ++q) s.c2[q] = (float)d->c2[3 * b + q];
    s.nonlin_std = d->nonlin_std[b], s.bf_std = d->bf_std[b], s.gamma = (float)d->gamma[b], s.noise_std = d->noise_std[b];
    s.tex = s.deform ? in->tex[b] : 0;
    s.surf = s.deform ? in->surf[b] : 0;
    for (int a = 0; a < 3; ++a) {
      const int64_t fs = d->size_f[3 * b + a], bs = d->bf_size[3 * b + a];
      s.fs[a] = (int32_t)fs, s.bs[a] = (int32_t)bs;
      if (s.deform) {
        if (fs < 1 || fs >= in->zoom_len[a] || !in->zoom_tab[a][fs]) return missing[0] = 0, missing[1] = a, missing[2] = (int32_t)fs, -2;
        s.ftab[a] = reinterpret_cast<const fsg_tab*>(in->zoom_tab[a][fs]);
      }
      if (s.bias_on) {
        if (bs < 1 || bs >= in->zoom_len[a] || !in->zoom_tab[a][bs]) return missing[0] = 0, missing[1] = a, missing[2] = (int32_t)bs, -2;
        s.btab[a] = reinterpret_cast<const fsg_tab*>(in->zoom_tab[a][bs]);
      }
    }
    if (s.res_on) {
      const double sp = d->spacing[b];
      const float* prev[3] = {nullptr, nullptr, nullptr};
      for (int a = 0; a < 3; ++a) {
        if (sp < c->res[a]) return -1;  // an up-sampled axis: generic path
        const int64_t n = (int64_t)((double)st->shape[a] * c->res[a] / sp);  // resample_size (synthseg.py:82-84)
        s.n_out[a] = (int32_t)n;
        if (n < 1 || n >= in->res_len[a] || !in->pos_tab[a][n] || !in->back_tab[a][n]) return missing[0] = 1, missing[1] = a, missing[2] = (int32_t)n, -2;
        s.pos[a] = reinterpret_cast<const fsg_tab*>(in->pos_tab[a][n]);
        s.ztab[a] = reinterpret_cast<const fsg_tab*>(in->back_tab[a][n]);
        s.ntaps[a] = 1;
        const double sd = d->stds[3 * b + a];
        if (sd > 0) {
          for (int p = 0; p < a; ++p)  // one tap array per distinct width of a sample
            if (d->stds[3 * b + p] == sd) s.taps[a] = prev[p], s.ntaps[a] = s.ntaps[p];
          if (!s.taps[a]) {
            float* t = in->taps_host + ((size_t)b * 3 + a) * FSG_STEP_MAX_TAPS;
            const int nt = fsg_gaussian_taps(sd, t, FSG_STEP_MAX_TAPS);
            if (nt < 0) return -1;
            s.taps[a] = t;
            s.ntaps[a] = nt;
          }
          prev[a] = s.taps[a];
        }
      }
    }
  }
  return 0;
}
