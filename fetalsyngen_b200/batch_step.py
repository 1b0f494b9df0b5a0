"""Launch builders of the batched base path: the native step (``run_step_native``: draws, sample structs, job structs
and launches in the library, csrc/step.cu) and its numpy fallback (``run_base_batch``).

``SynthEngine.run_base`` fills one ctypes job struct per sample and stage, field by field (~250 attribute
stores and as many small conversions per step of 8 samples: 0.9 ms of Python).  Here the same job arrays are
numpy structured arrays over the batch (``_lib.np_dtype`` mirrors the C structs), filled column by column
from the arrays ``batch_draw.draw_batch`` draws, with buffer addresses computed from base pointers instead
of per-row tensor views.  The launches, their order and every job field are the ones ``_run_base`` produces
(``tests/test_host.py`` compares the job bytes of the two builders; the GPU suite compares their outputs),
so this is purely a cheaper way to issue the same work.

Covered: the production configuration — Philox noise, control grids drawn on the device, seeds as packed
words or as label volumes, any combination of the per-sample gates (deformation, flip, gamma, bias field,
resolution simulation, noise).  Anything else (injected tensors, a second image channel, a simulated spacing
finer than the input resolution, extents the texture hand-over does not take) returns False and the caller
takes the generic path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .engine import STAGE_BIAS, STAGE_FIELD, STAGE_GMM, STAGE_NOISE, _PAIRS, _stream
from .tables import gaussian_taps_np

_U64 = np.uint64


def _dt(name):
    return _lib.np_dtype(getattr(_lib, name))


def _as_arg(arr: np.ndarray, ctype):
    return C.cast(arr.ctypes.data, C.POINTER(ctype))


def _rows(eng, name, B, dtype=torch.float32, numel=None):
    """(address of row 0, row pitch in bytes) of an engine scratch buffer."""
    t = eng.scratch(name, B, dtype, numel)
    return t.data_ptr(), t.stride(0) * t.element_size()


def run_base_batch(eng, d, seeds, segs, out_img, out_seg, scale) -> bool:
    B = d.B
    sx, sy, sz = eng.shape
    nvox = eng.nvox
    res = eng.resolution
    if _PAIRS or not eng.use_tex or sx % 8 or sy % 4 or sz % 4 or min(eng.shape) < 2 or B > _lib.MAX_JOBS:
        return False
    if d.deform_on.any() and not d.nonlinear:
        return False
    rs = np.flatnonzero(d.res_on)
    if rs.size and (d.spacing[rs, None] < res[None, :]).any():
        return False  # an up-sampled axis: the coarse grid outgrows the volume (generic path sizes its scratch for it)
    idx = np.arange(B)
    dev = eng.device

    # ------------------------------------------------------------------ buffers
    buf0, p0 = _rows(eng, "buf0", B)
    buf1, p1 = _rows(eng, "buf1", B)
    buf2, p2 = _rows(eng, "buf2", B)
    oimg, oseg = out_img.data_ptr(), out_seg.data_ptr()
    seg_ptr = np.fromiter((s.data_ptr() for s in segs), dtype=_U64, count=B)
    for s in segs:
        if s.numel() != nvox or s.dtype != torch.uint8 or s.device != dev or not s.is_contiguous():
            return False
    deform = d.deform_on
    tex_on = deform.copy()  # deformation with a control grid on tile-aligned extents: texture hand-over
    texv = [eng.texvol(b) if tex_on[b] else None for b in range(B)]

    eng.begin()
    try:
        _issue(eng, d, seeds, B, idx, rs, deform, tex_on, texv, seg_ptr, (buf0, p0), (buf1, p1), (buf2, p2), oimg, oseg, scale, out_img)
    except BaseException:
        eng._batch = None
        eng.tables.hold = False
        raise
    eng.flush()
    return True


def _ring_block(eng, nfloats):
    """Reserve `nfloats` (multiple of 4) in the open batch's parameter ring: (host float32 view, device address)."""
    b = eng._batch
    if b["used"] + nfloats > eng.RING_FLOATS:
        raise RuntimeError("parameter ring overflow")
    o = b["used"]
    b["used"] = o + nfloats
    return eng._ring_np[b["slot"]][o : o + nfloats], eng._ring[1][b["slot"]].data_ptr() + 4 * o


def _issue(eng, d, seeds, B, idx, rs, deform, tex_on, texv, seg_ptr, b0, b1, b2, oimg, oseg, scale, out_img):
    sx, sy, sz = eng.shape
    nvox = eng.nvox
    res = eng.resolution
    keep = eng._batch["keep"]
    seed_key = _U64(d.base_seed & (2**64 - 1))
    buf0, p0 = b0
    buf1, p1 = b1
    buf2, p2 = b2
    row0 = (buf0 + idx * p0).astype(_U64)
    row1 = (buf1 + idx * p1).astype(_U64)
    row2 = (buf2 + idx * p2).astype(_U64)
    out_rows = (oimg + idx * (4 * nvox)).astype(_U64)
    seg_rows = (oseg + idx * nvox).astype(_U64)

    # ------------------------------------------------------------------ K1: GMM
    nl = d.mus.shape[1]
    nlp = (nl + 3) // 4 * 4
    hv, dptr = _ring_block(eng, 2 * B * nlp)
    hv = hv.reshape(2, B, nlp)
    hv[0, :, :nl] = d.mus
    hv[1, :, :nl] = d.sigmas
    g = np.zeros(B, dtype=_dt("GmmJob"))
    g["mus"] = dptr + 4 * nlp * idx
    g["sigmas"] = dptr + 4 * nlp * (B + idx)
    g["nlabels"] = nl
    g["rng"]["seed"], g["rng"]["sample"], g["rng"]["stage"] = seed_key, d.sample_ids, STAGE_GMM
    kinds = np.zeros(B, dtype=np.int64)
    for b, sd in enumerate(seeds):
        if isinstance(sd, tuple):
            ps, m2s = sd
            if int(np.prod(ps.shape)) != nvox:
                raise ValueError("packed seed words must hold one word per output voxel")
            for m in range(1, 5):
                n = int(m2s[m])
                if n not in ps.layout:
                    raise KeyError(f"no seeds with {n} sub-classes in this cache (available: {ps.counts})")
                g["shift"][b, m - 1], g["mask"][b, m - 1] = ps.layout[n]
            g["words"][b], g["word_bytes"][b] = ps.on(eng.device).data_ptr(), ps.word_bytes
        else:
            vols = list(sd)
            if not 1 <= len(vols) <= 4:
                raise ValueError("each sample needs 1..4 seed volumes")
            for m, v in enumerate(vols):
                if v.dtype not in (torch.int8, torch.uint8) or v.numel() != nvox or v.device != eng.device or not v.is_contiguous():
                    raise TypeError("seed volumes must be contiguous int8/uint8 device tensors with one label per output voxel")
                g["seed"][b, m] = v.data_ptr()
            kinds[b] = len(vols)
    g["out"] = np.where(tex_on, 0, row0)
    g["out_surf"] = [0 if t is None else t.h.surf for t in texv]
    g["row_len"] = np.where(tex_on, sz, 0)
    g["surf_ny"] = np.where(tex_on, sy, 0)
    keep.append(g)
    for k in np.unique(kinds):  # a launch needs one kind of label source
        sel = np.flatnonzero(kinds == k)
        sub = g if sel.size == B else np.ascontiguousarray(g[sel])
        keep.append(sub)
        eng._call("fsg_gmm", _as_arg(sub, _lib.GmmJob), int(sel.size), nvox)

    # ------------------------------------------------------------------ K2: control grids + warp
    bias = d.bias_on
    nf_each = 3 * d.size_f.prod(axis=1)
    nb_each = d.bf_size.prod(axis=1)
    nf = (int(nf_each[deform].max(initial=0)) + 3) // 4 * 4
    nb = (int(nb_each[bias].max(initial=0)) + 3) // 4 * 4
    ngrid = int(deform.sum() + bias.sum())
    grows = 0
    if ngrid:
        gbuf, gp = _rows(eng, "grids", B, torch.float32, max(nf + nb, 4))
        grows = (gbuf + idx * gp).astype(_U64)
        gj = np.zeros(ngrid, dtype=_dt("GridJob"))
        # job order of the generic builder: per sample, field grid then bias grid
        order = np.concatenate([np.stack([idx[deform], np.zeros(int(deform.sum()), dtype=np.int64)], 1), np.stack([idx[bias], np.ones(int(bias.sum()), dtype=np.int64)], 1)])
        order = order[np.lexsort((order[:, 1], order[:, 0]))]
        ob, ok = order[:, 0], order[:, 1].astype(bool)
        gj["out"] = grows[ob] + np.where(ok, 4 * nf, 0).astype(_U64)
        gj["n"] = np.where(ok, nb_each[ob], nf_each[ob])
        gj["scale"] = np.where(ok, d.bf_std[ob], d.nonlin_std[ob])
        gj["rng"]["seed"], gj["rng"]["sample"], gj["rng"]["stage"] = seed_key, d.sample_ids[ob], np.where(ok, STAGE_BIAS, STAGE_FIELD)
        keep.append(gj)
        for q0 in range(0, ngrid, _lib.MAX_JOBS):
            n = min(_lib.MAX_JOBS, ngrid - q0)
            eng._call("fsg_draw_grids", C.cast(gj.ctypes.data + q0 * gj.dtype.itemsize, C.POINTER(_lib.GridJob)), n)
    shift, sp = _rows(eng, "shift", B, torch.float32, 4)
    w = np.zeros(B, dtype=_dt("WarpJob"))
    w["src_img"] = np.where(tex_on, 0, row0)
    w["src_tex"] = [0 if t is None else t.h.tex for t in texv]
    w["src_seg"], w["dst_seg"] = seg_ptr, seg_rows
    res_on, noise_on = d.res_on, d.noise_on
    direct = ~res_on & ~noise_on  # warp straight into the output
    w["dst_img"] = np.where(direct, out_rows, row1)
    w["mode"] = deform
    w["flip"] = deform & d.flip
    w["shift"] = shift + idx * sp
    if deform.any():
        w["A"][deform] = d.A.reshape(B, 9)[deform]
        w["c2"][deform] = d.c2[deform].astype(np.float32)
        w["center"][deform] = np.asarray(d.center, dtype=np.float32)[None, :]
        w["fsmall"][deform] = grows[deform]
        w["fs"][deform] = d.size_f[deform]
        for a in range(3):
            w["ftab"][deform, a] = [eng.zoom_table_ptr(int(n), a) for n in d.size_f[deform, a]]
    g_on = d.gamma_on
    w["has_gamma"] = g_on
    w["gamma"] = np.where(g_on, d.gamma, 0.0).astype(np.float32)
    if bias.any():
        w["bf_low"][bias] = grows[bias] + _U64(4 * nf)
        w["bs"][bias] = d.bf_size[bias]
        for a in range(3):
            w["btab"][bias, a] = [eng.zoom_table_ptr(int(n), a) for n in d.bf_size[bias, a]]
    keep.append(w)
    if deform.any():
        dj = w if deform.all() else np.ascontiguousarray(w[deform])
        keep.append(dj)
        eng._call("fsg_warp_shift", _as_arg(dj, _lib.WarpJob), int(deform.sum()), sx, sy, sz)
    eng._call("fsg_warp", _as_arg(w, _lib.WarpJob), B, sx, sy, sz)

    # ------------------------------------------------------------------ K4: resolution simulation + zoom back
    R = int(rs.size)
    if R:
        ridx = np.arange(R)
        spc = d.spacing[rs]
        n_out = np.stack([(eng.shape[a] * res[a] / spc).astype(np.int64) for a in range(3)], 1)  # resample_size
        taps, tptr, ntaps = [], np.zeros((R, 3), dtype=_U64), np.ones((R, 3), dtype=np.int64)
        for k, b in enumerate(rs):  # Gaussian taps: one array per distinct width of a sample (continuous: not cached)
            seen = {}
            for a in range(3):
                sdv = float(d.stds[b, a])
                if sdv > 0:
                    if sdv not in seen:
                        seen[sdv] = len(taps)
                        taps.append(gaussian_taps_np(sdv))
                    tptr[k, a], ntaps[k, a] = seen[sdv], taps[seen[sdv]].size
                else:
                    tptr[k, a] = _U64(2**63)  # no taps
        if taps:
            sizes = np.array([t.size for t in taps])
            offs = np.concatenate([[0], np.cumsum((sizes + 3) // 4 * 4)])
            hv, dptr = _ring_block(eng, int(offs[-1]))
            for t, o in zip(taps, offs):
                hv[o : o + t.size] = t
            has = tptr < _U64(2**63)
            tptr = np.where(has, dptr + 4 * offs[np.where(has, tptr, 0).astype(np.int64)], 0).astype(_U64)
            maxw = max(2, int(sizes.max()) + 1)
        else:
            tptr[:] = 0
            maxw = 2
        nmax = max(max(eng.shape), int(n_out.max()))
        maxw = max(32, (maxw + 3) // 4 * 4)
        per_axis = (nmax * maxw + (nmax + 1) // 2 + 3) // 4 * 4
        ws, wp = _rows(eng, "sep_tables", R, torch.float32, 3 * per_axis)
        n_in = np.asarray(eng.shape, dtype=np.int64)[None, :].repeat(R, 0)
        width = np.minimum(n_in, ntaps + 1)
        w_ptr = (ws + ridx[:, None] * wp + 4 * per_axis * np.arange(3)[None, :]).astype(_U64)
        q_ptr = w_ptr + _U64(4 * nmax * maxw)
        pos = np.zeros((R, 3), dtype=_U64)
        fac = np.zeros((R, 3))
        for a in range(3):
            pf = [eng.resample_table_ptr(a, float(s)) for s in spc]
            pos[:, a] = [p[0] for p in pf]
            fac[:, a] = [p[1] for p in pf]
        cj = np.zeros((R, 3), dtype=_dt("SepComposeJob"))
        cj["pos"], cj["taps"], cj["q0_out"], cj["w_out"] = pos, tptr, q_ptr, w_ptr
        cj["ntaps"], cj["n_in"], cj["n_out"], cj["width"] = ntaps, n_in, n_out, width
        cj["cap_q0"], cj["cap_w"] = nmax, nmax * maxw
        sj = np.zeros(R, dtype=_dt("SepconvJob"))
        ax = sj["ax"]
        ax["q0"], ax["w"], ax["n_out"], ax["width"], ax["pos"], ax["taps"], ax["ntaps"] = q_ptr, w_ptr, n_out, width, pos, tptr, ntaps
        sj["src"], sj["dst"], sj["tmp1"], sj["tmp2"] = row1[rs], row2[rs], row2[rs], row0[rs]
        sj["cap_dst"] = sj["cap_tmp1"] = sj["cap_tmp2"] = nvox
        nz = noise_on[rs]
        sj["has_noise"] = nz
        sj["noise_std"] = np.where(nz, d.noise_std[rs], 0).astype(np.float32)
        sj["rng"]["seed"] = np.where(nz, seed_key, 0).astype(_U64)
        sj["rng"]["sample"] = np.where(nz, d.sample_ids[rs], 0).astype(_U64)
        sj["rng"]["stage"] = np.where(nz, STAGE_NOISE, 0)
        keep.extend([cj, sj])
        eng._call("fsg_sep_compose", _as_arg(cj, _lib.SepComposeJob), 3 * R)
        eng._call("fsg_sepconv", _as_arg(sj, _lib.SepconvJob), R, sx, sy, sz)
        zj = np.zeros(R, dtype=_dt("ZoomJob"))
        mm, mp = _rows(eng, "minmax", R, torch.float32, 2)
        for a in range(3):
            zj["tab"][:, a] = [eng.zoom_back_ptr(a, int(n), float(f)) for n, f in zip(n_out[:, a], fac[:, a])]
        zj["n"], zj["src"], zj["dst"] = n_out, row2[rs], out_rows[rs]
        zj["minmax"], zj["post"] = mm + ridx * mp, 2 if scale else 1
        keep.append(zj)
        eng._call("fsg_zoom_minmax", _as_arg(zj, _lib.ZoomJob), R, sx, sy, sz)
        eng._call("fsg_zoom", _as_arg(zj, _lib.ZoomJob), R, sx, sy, sz)

    # ------------------------------------------------------------------ samples without the resolution simulation
    nzo = np.flatnonzero(~res_on & noise_on)
    if nzo.size:
        nj = np.zeros(nzo.size, dtype=_dt("NoiseJob"))
        nj["src"], nj["dst"], nj["noise_std"] = row1[nzo], out_rows[nzo], d.noise_std[nzo]
        nj["rng"]["seed"], nj["rng"]["sample"], nj["rng"]["stage"] = seed_key, d.sample_ids[nzo], STAGE_NOISE
        keep.append(nj)
        eng._call("fsg_add_noise", _as_arg(nj, _lib.NoiseJob), int(nzo.size), nvox)
    if scale:
        for b in np.flatnonzero(~res_on):
            eng.scale_intensity(out_img[b], out_img[b])


# ---------------------------------------------------------------------------------------------- native builder
def _lookup(eng, kind, axis, keys, make):
    """Addresses of cached 1-D tables for an integer key per sample: a dense per-engine array indexed by the key
    (one fancy-indexing operation per step), filled through `make(key)` on first use."""
    tab = eng._ptrs.get((kind, axis))
    if tab is None:
        tab = eng._ptrs[(kind, axis)] = np.zeros(eng.shape[axis] + 2, dtype=_U64)
    out = tab.take(keys, mode="clip")
    if not out.all():
        keys = np.minimum(np.maximum(keys, 0), tab.size - 1)
        for k in np.unique(keys[out == 0]):
            tab[k] = make(int(k))
        out = tab[keys]
    return out


def fill_step(eng, d, seeds, segs, out_img, out_seg, scale, keep):
    """(fsg_step, fsg_step_sample[B]) of one batch for fsg_step_run / fsg_step_build, or None when the step is not
    one the native builder covers.  `keep` receives the host arrays the structs point into."""
    B = d.B
    sx, sy, sz = eng.shape
    nvox = eng.nvox
    res = eng.resolution
    if _PAIRS or not eng.use_tex or sx % 8 or sy % 4 or sz % 4 or min(eng.shape) < 2 or B > _lib.MAX_JOBS:
        return None
    if d.deform_on.any() and not d.nonlinear:
        return None
    rs = np.flatnonzero(d.res_on)
    if rs.size and (d.spacing[rs, None] < res[None, :]).any():
        return None
    for s in segs:
        if s.numel() != nvox or s.dtype != torch.uint8 or s.device != eng.device or not s.is_contiguous():
            return None
    idx = np.arange(B)
    S = np.zeros(B, dtype=_dt("StepSample"))
    S["seg"] = np.fromiter((s.data_ptr() for s in segs), dtype=_U64, count=B)
    for b, sd in enumerate(seeds):
        if isinstance(sd, tuple):
            ps, m2s = sd
            if int(np.prod(ps.shape)) != nvox:
                raise ValueError("packed seed words must hold one word per output voxel")
            for m in range(1, 5):
                n = int(m2s[m])
                if n not in ps.layout:
                    raise KeyError(f"no seeds with {n} sub-classes in this cache (available: {ps.counts})")
                S["shift"][b, m - 1], S["mask"][b, m - 1] = ps.layout[n]
            S["words"][b], S["word_bytes"][b] = ps.on(eng.device).data_ptr(), ps.word_bytes
        else:
            vols = list(sd)
            if not 1 <= len(vols) <= 4:
                raise ValueError("each sample needs 1..4 seed volumes")
            for m, v in enumerate(vols):
                if v.dtype not in (torch.int8, torch.uint8) or v.numel() != nvox or v.device != eng.device or not v.is_contiguous():
                    raise TypeError("seed volumes must be contiguous int8/uint8 device tensors with one label per output voxel")
                S["seed"][b, m] = v.data_ptr()
    deform, bias = d.deform_on, d.bias_on
    S["deform"], S["flip"], S["gamma_on"], S["bias_on"], S["res_on"], S["noise_on"] = deform, d.flip, d.gamma_on, bias, d.res_on, d.noise_on
    S["sample_id"] = d.sample_ids
    mus, sigmas = np.ascontiguousarray(d.mus, dtype=np.float32), np.ascontiguousarray(d.sigmas, dtype=np.float32)
    keep.extend([mus, sigmas, S])
    nl = mus.shape[1]
    S["mus"] = mus.ctypes.data + 4 * nl * idx
    S["sigmas"] = sigmas.ctypes.data + 4 * nl * idx
    S["A"] = d.A.reshape(B, 9)
    S["c2"] = d.c2.astype(np.float32)
    S["nonlin_std"], S["bf_std"], S["gamma"], S["noise_std"] = d.nonlin_std, d.bf_std, d.gamma.astype(np.float32), d.noise_std
    S["fs"], S["bs"] = d.size_f, d.bf_size
    texv = [eng.texvol(b) if deform[b] else None for b in range(B)]
    S["tex"] = [0 if t is None else t.h.tex for t in texv]
    S["surf"] = [0 if t is None else t.h.surf for t in texv]
    for a in range(3):
        S["ftab"][:, a] = _lookup(eng, "z", a, d.size_f[:, a], lambda n, a=a: eng.zoom_table_ptr(n, a))
        S["btab"][:, a] = _lookup(eng, "z", a, d.bf_size[:, a], lambda n, a=a: eng.zoom_table_ptr(n, a))
    maxw = 2
    if rs.size:
        n_out = (np.asarray(eng.shape, dtype=np.float64)[None, :] * res[None, :] / d.spacing[:, None]).astype(np.int64)  # resample_size
        S["n_out"] = n_out
        for a in range(3):
            def make_pos(n, a=a):
                b = int(np.flatnonzero(n_out[:, a] == n)[0])
                return eng.resample_table_ptr(a, float(d.spacing[b]))[0]

            def make_back(n, a=a):
                b = int(np.flatnonzero(n_out[:, a] == n)[0])
                return eng.zoom_back_ptr(a, n, eng.resample_table_ptr(a, float(d.spacing[b]))[1])

            S["pos"][:, a] = _lookup(eng, "r", a, n_out[:, a], make_pos)
            S["ztab"][:, a] = _lookup(eng, "b", a, n_out[:, a], make_back)
        S["ntaps"] = 1
        for b in rs:  # Gaussian taps: one array per distinct width of a sample (continuous widths: not cached)
            seen = {}
            for a in range(3):
                sdv = float(d.stds[b, a])
                if sdv > 0:
                    t = seen.get(sdv)
                    if t is None:
                        t = seen[sdv] = gaussian_taps_np(sdv)
                        keep.append(t)
                        maxw = max(maxw, t.size + 1)
                    S["taps"][b, a], S["ntaps"][b, a] = t.ctypes.data, t.size
    # ---- engine buffers
    st = _lib.Step()
    st.B, st.nlabels, st.scale, st.seed = B, nl, int(bool(scale)), d.base_seed & (2**64 - 1)
    st.shape = (C.c_int32 * 3)(sx, sy, sz)
    st.center = (C.c_float * 3)(*np.asarray(d.center, dtype=np.float32))
    for k, name in enumerate(("buf0", "buf1", "buf2")):
        st.buf[k], st.buf_pitch[k] = _rows(eng, name, B)
    st.out_img, st.out_seg = out_img.data_ptr(), out_seg.data_ptr()
    nf = (int((3 * d.size_f.prod(axis=1))[deform].max(initial=0)) + 3) // 4 * 4
    nb = (int(d.bf_size.prod(axis=1)[bias].max(initial=0)) + 3) // 4 * 4
    cap = max(nf + nb, 4)
    st.grids, st.grids_pitch = _rows(eng, "grids", B, torch.float32, cap)
    st.grids_cap = cap
    st.shift, st.shift_pitch = _rows(eng, "shift", B, torch.float32, 4)
    if rs.size:
        nmax = max(max(eng.shape), int(S["n_out"][rs].max()))
        mw = max(32, (maxw + 3) // 4 * 4)
        per_axis = (nmax * mw + (nmax + 1) // 2 + 3) // 4 * 4
        st.sep_tables, st.sep_pitch = _rows(eng, "sep_tables", int(rs.size), torch.float32, 3 * per_axis)
        st.sep_cap = 3 * per_axis
    st.minmax, st.minmax_pitch = _rows(eng, "minmax", B, torch.float32, 2)
    return st, S


# ---------------------------------------------------------------------------------------------- fully native step
class _NativeState:
    """Per-engine host arrays of the native step: the inputs struct with its arrays, the dense table-address arrays,
    the sample structs and the step struct (buffer fields refreshed per call: scratch may have been regrown)."""

    def __init__(self, eng):
        from .data.packed import PackedSeeds  # noqa: F401  (layout arrays are cached on the PackedSeeds objects)

        M = _lib.MAX_JOBS
        self.seg = np.zeros(M, dtype=_U64)
        self.words = np.zeros(M, dtype=_U64)
        self.word_bytes = np.zeros(M, dtype=np.int32)
        self.layout = np.zeros(M, dtype=_U64)
        self.layout_len = np.zeros(M, dtype=np.int32)
        self.seed = np.zeros((M, 4), dtype=_U64)
        self.tex = np.zeros(M, dtype=_U64)
        self.surf = np.zeros(M, dtype=_U64)
        self.taps = np.zeros((M, 3, _lib.STEP_MAX_TAPS), dtype=np.float32)
        self.zoom = [np.zeros(eng.shape[a] + 2, dtype=_U64) for a in range(3)]
        self.pos = [np.zeros(eng.shape[a] + 2, dtype=_U64) for a in range(3)]
        self.back = [np.zeros(eng.shape[a] + 2, dtype=_U64) for a in range(3)]
        self.missing = np.zeros(4, dtype=np.int32)
        self.S = (_lib.StepSample * M)()
        i = self.inputs = _lib.StepInputs()
        i.seg, i.words, i.word_bytes, i.layout, i.layout_len, i.seed = (a.ctypes.data for a in (self.seg, self.words, self.word_bytes, self.layout, self.layout_len, self.seed))
        i.tex, i.surf, i.taps_host = self.tex.ctypes.data, self.surf.ctypes.data, self.taps.ctypes.data
        for a in range(3):
            i.zoom_tab[a], i.pos_tab[a], i.back_tab[a] = self.zoom[a].ctypes.data, self.pos[a].ctypes.data, self.back[a].ctypes.data
            i.zoom_len[a] = i.res_len[a] = eng.shape[a] + 2
        self.st = _lib.Step()
        self.st.shape = (C.c_int32 * 3)(*eng.shape)
        self.ntex = 0


def _layout_array(ps):
    """[nmax + 1][2] int32 (shift, mask) per sub-class count of a packed subject (-1: count absent), cached."""
    arr = getattr(ps, "_layout_arr", None)
    if arr is None:
        nmax = max(ps.layout)
        arr = np.full((nmax + 1, 2), -1, dtype=np.int32)
        for n, (sh, mk) in ps.layout.items():
            arr[n] = (sh, mk)
        ps._layout_arr = arr
        ps._nvox = int(np.prod(ps.shape))
    return arr


def prepare_step_native(eng, gen, d, seeds, segs, out_img, out_seg, scale):
    """(fsg_step, fsg_step_sample array) of one batch, filled by the library (`fsg_step_fill`) from the draws and the
    input addresses gathered here; None when the step is not covered by the native path."""
    from .batch_draw import _draw_config
    from .data.packed import PackedSeeds

    B = d.B
    sx, sy, sz = eng.shape
    nvox = eng.nvox
    if _PAIRS or not eng.use_tex or sx % 8 or sy % 4 or sz % 4 or min(eng.shape) < 2 or B > _lib.MAX_JOBS or len(seeds) != B or len(segs) != B:
        return None
    ns = eng.__dict__.get("_native")
    if ns is None:
        ns = eng._native = _NativeState(eng)
    dev = eng.device
    for b in range(B):
        s, sd = segs[b], seeds[b]
        if s.numel() != nvox or s.dtype != torch.uint8 or s.device != dev or not s.is_contiguous():
            return None
        ns.seg[b] = s.data_ptr()
        if isinstance(sd, PackedSeeds):
            la = _layout_array(sd)
            if sd._nvox != nvox or d.m2s is None:
                return None
            ns.words[b], ns.word_bytes[b], ns.layout[b], ns.layout_len[b] = sd.on(dev).data_ptr(), sd.word_bytes, la.ctypes.data, la.shape[0]
        elif isinstance(sd, (list, tuple)) and 1 <= len(sd) <= 4 and all(torch.is_tensor(v) for v in sd):
            ns.words[b] = 0
            ns.seed[b] = 0
            for m, v in enumerate(sd):
                if v.dtype not in (torch.int8, torch.uint8) or v.numel() != nvox or v.device != dev or not v.is_contiguous():
                    return None
                ns.seed[b, m] = v.data_ptr()
        else:
            return None
    while ns.ntex < B:  # the engine's block-linear volumes, one per batch slot
        t = eng.texvol(ns.ntex)
        ns.tex[ns.ntex], ns.surf[ns.ntex] = t.h.tex, t.h.surf
        ns.ntex += 1
    cfg = _draw_config(gen, tuple(eng.shape))[0]
    st = ns.st
    st.B, st.nlabels, st.scale, st.seed = B, cfg.nlabels, int(bool(scale)), d.base_seed & (2**64 - 1)
    st.center = (C.c_float * 3)(*d.center)
    for k, name in enumerate(("buf0", "buf1", "buf2")):
        st.buf[k], st.buf_pitch[k] = _rows(eng, name, B)
    st.out_img, st.out_seg = out_img.data_ptr(), out_seg.data_ptr()
    caps = eng.__dict__.get("_native_caps")
    if caps is None:  # worst-case scratch sizes of this generator's ranges
        shp = np.asarray(eng.shape, dtype=np.float64)
        nf = 3 * int(np.prod(np.round(cfg.nonlin_scale_max * shp) + 1))
        nb = int(np.prod(np.maximum(np.round(cfg.bf_scale_max * shp), 1) + 1))
        nmax, mw = max(eng.shape), _lib.STEP_MAX_TAPS + 4
        caps = eng._native_caps = ((nf + 3) // 4 * 4 + (nb + 3) // 4 * 4, 3 * ((nmax * mw + (nmax + 1) // 2 + 3) // 4 * 4))
    st.grids, st.grids_pitch = _rows(eng, "grids", B, torch.float32, caps[0])
    st.grids_cap = caps[0]
    st.shift, st.shift_pitch = _rows(eng, "shift", B, torch.float32, 4)
    st.sep_tables, st.sep_pitch = _rows(eng, "sep_tables", B, torch.float32, caps[1])
    st.sep_cap = caps[1]
    st.minmax, st.minmax_pitch = _rows(eng, "minmax", B, torch.float32, 2)
    lib = _lib.load()
    out = d._c_out
    for _ in range(64):
        rc = lib.fsg_step_fill(C.byref(st), C.byref(cfg), C.byref(out), d.sample_ids.ctypes.data, C.byref(ns.inputs), ns.S, ns.missing.ctypes.data)
        if rc != -2:
            break
        kind, a, key = (int(v) for v in ns.missing[:3])  # build the missing table, then fill again
        if kind == 0:
            ns.zoom[a][key] = eng.zoom_table_ptr(key, a)
        else:
            b = next(b for b in range(B) if d.res_on[b] and int(eng.shape[a] * eng.resolution[a] / d.spacing[b]) == key)
            p, fac = eng.resample_table_ptr(a, float(d.spacing[b]))
            ns.pos[a][key], ns.back[a][key] = p, eng.zoom_back_ptr(a, key, fac)
    if rc == -1:
        return None
    if rc:
        raise _lib.FsgError(f"fsg_step_fill failed ({rc}): {lib.fsg_last_error().decode()}")
    return st, ns.S


def run_step_native(eng, gen, d, seeds, segs, out_img, out_seg, scale) -> bool:
    """One batched step with everything after the draws in the library: `fsg_step_fill` turns the draws (made by
    `fsg_draw_batch`) and the input addresses into sample structs, `fsg_step_run` builds the jobs and launches.
    Python only gathers addresses.  Returns False when the step is not covered (callers fall back)."""
    prepared = prepare_step_native(eng, gen, d, seeds, segs, out_img, out_seg, scale)
    if prepared is None:
        return False
    st, S = prepared
    lib = _lib.load()
    eng.begin()
    bt = eng._batch
    try:
        hring, dring, events, _ = eng._ring
        st.ring_host, st.ring_dev, st.ring_floats = hring[bt["slot"]].data_ptr(), dring[bt["slot"]].data_ptr(), eng.RING_FLOATS
        rc = lib.fsg_step_run(C.byref(st), S, _stream())
        if rc > 0:
            raise _lib.FsgError(f"fsg_step_run failed ({rc}): {lib.fsg_last_error().decode()}")
        if rc == 0:
            _lib.stats.calls["fsg_step_run"] = _lib.stats.calls.get("fsg_step_run", 0) + 1
            e = torch.cuda.Event()
            e.record()
            events[bt["slot"]] = e
    finally:
        eng._batch = None
        eng.tables.hold = False
    eng._keep_native = d  # the sample structs point into the draw's arrays until the next step replaces them
    return rc == 0
