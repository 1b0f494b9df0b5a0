"""ctypes binding of libfsg.so (the C-ABI declared in include/fsg.h).

The product path has no CPU fallback: if the library is missing or a launch fails, a
``RuntimeError`` is raised.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

PKG = Path(__file__).resolve().parent
import os

LIB_PATH = Path(os.environ.get("FSG_LIB", PKG / "libfsg.so"))  # FSG_LIB: A/B builds of the kernels (benchmarks only)
MAX_JOBS = 16
MAX_TAPS = 127

_vp, _i32, _i64, _f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float


class Tab(C.Structure):
    _fields_ = [("f", C.c_int16), ("c", C.c_int16), ("wc", _f32)]


class Rng(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("sample", C.c_uint64), ("stage", C.c_uint32), ("_pad", C.c_uint32)]


class GmmJob(C.Structure):
    _fields_ = [("seed", _vp * 4), ("mus", _vp), ("sigmas", _vp), ("noise", _vp), ("out", _vp), ("labels_out", _vp),
                ("rng", Rng), ("nlabels", _i32), ("row_len", _i32), ("out_pairs", _vp), ("pairs_float", _i32), ("word_bytes", _i32), ("words", _vp), ("shift", _i32 * 4), ("mask", _i32 * 4),
                ("out_surf", C.c_uint64), ("surf_ny", _i32), ("_pad2", _i32)]


class WarpJob(C.Structure):
    _fields_ = [("src_img", _vp), ("src_pairs", _vp), ("src_seg", _vp), ("src_img2", _vp), ("dst_img", _vp), ("dst_seg", _vp), ("dst_img2", _vp),
                ("fsmall", _vp), ("ftab", _vp * 3), ("bf_low", _vp), ("btab", _vp * 3), ("shift", _vp),
                ("A", _f32 * 9), ("c2", _f32 * 3), ("center", _f32 * 3), ("gamma", _f32),
                ("fs", _i32 * 3), ("bs", _i32 * 3), ("mode", _i32), ("flip", _i32), ("has_gamma", _i32), ("pairs_float", _i32),
                ("src_tex", C.c_uint64)]


class StepSample(C.Structure):  # struct fsg_step_sample
    _fields_ = [("seg", _vp), ("words", _vp), ("seed", _vp * 4), ("word_bytes", _i32), ("shift", _i32 * 4), ("mask", _i32 * 4),
                ("deform", _i32), ("flip", _i32), ("gamma_on", _i32), ("bias_on", _i32), ("res_on", _i32), ("noise_on", _i32),
                ("sample_id", C.c_uint64), ("mus", _vp), ("sigmas", _vp), ("A", _f32 * 9), ("c2", _f32 * 3),
                ("nonlin_std", _f32), ("bf_std", _f32), ("gamma", _f32), ("noise_std", _f32), ("fs", _i32 * 3), ("bs", _i32 * 3),
                ("tex", C.c_uint64), ("surf", C.c_uint64), ("ftab", _vp * 3), ("btab", _vp * 3), ("pos", _vp * 3), ("ztab", _vp * 3),
                ("n_out", _i32 * 3), ("ntaps", _i32 * 3), ("taps", _vp * 3)]


class Step(C.Structure):  # struct fsg_step
    _fields_ = [("B", _i32), ("nlabels", _i32), ("shape", _i32 * 3), ("scale", _i32), ("center", _f32 * 3), ("_pad", _i32), ("seed", C.c_uint64),
                ("buf", _vp * 3), ("buf_pitch", _i64 * 3), ("out_img", _vp), ("out_seg", _vp), ("grids", _vp), ("grids_pitch", _i64), ("grids_cap", _i64),
                ("shift", _vp), ("shift_pitch", _i64), ("sep_tables", _vp), ("sep_pitch", _i64), ("sep_cap", _i64), ("minmax", _vp), ("minmax_pitch", _i64),
                ("ring_host", _vp), ("ring_dev", _vp), ("ring_floats", _i64)]


_f64, _u8p = C.c_double, C.c_void_p


class DrawConfig(C.Structure):  # struct fsg_draw_config
    _fields_ = [("nlabels", _i32), ("nseed", _i32), ("seed_labels", _vp), ("generation_classes", _vp), ("tied", _i32), ("meta_labels", _i32),
                ("min_subclusters", _i32), ("max_subclusters", _i32), ("shape", _i32 * 3), ("nonlinear", _i32), ("res", _f64 * 3),
                ("deform_prob", _f64), ("flip_prb", _f64), ("max_rotation", _f64), ("max_shear", _f64), ("max_scaling", _f64), ("nonlin_scale_min", _f64),
                ("nonlin_scale_max", _f64), ("nonlin_std_max", _f64), ("centre2", _f64 * 3), ("max_shift", _f64 * 3), ("gamma_prob", _f64), ("gamma_std", _f64),
                ("bias_prob", _f64), ("bf_scale_min", _f64), ("bf_scale_max", _f64), ("bf_std_min", _f64), ("bf_std_max", _f64),
                ("res_prob", _f64), ("min_resolution", _f64), ("max_resolution", _f64), ("noise_prob", _f64), ("noise_std_min", _f64), ("noise_std_max", _f64)]


class DrawOut(C.Structure):  # struct fsg_draw_out
    _fields_ = [(n, _vp) for n in ("mus", "sigmas", "deform_on", "flip", "gamma_on", "bias_on", "res_on", "noise_on", "rot", "shear", "scal", "A", "c2",
                                   "nonlin_scale", "size_f", "nonlin_std", "gamma", "bf_scale", "bf_size", "bf_std", "spacing", "stds", "noise_std", "m2s")]


class StepInputs(C.Structure):  # struct fsg_step_inputs
    _fields_ = [("seg", _vp), ("words", _vp), ("word_bytes", _vp), ("layout", _vp), ("layout_len", _vp), ("seed", _vp), ("tex", _vp), ("surf", _vp),
                ("zoom_tab", _vp * 3), ("zoom_len", _i32 * 3), ("pos_tab", _vp * 3), ("back_tab", _vp * 3), ("res_len", _i32 * 3), ("taps_host", _vp)]


STEP_MAX_TAPS = 64


class TexVol(C.Structure):
    _fields_ = [("array", C.c_uint64), ("tex", C.c_uint64), ("surf", C.c_uint64), ("nx", _i32), ("ny", _i32), ("nz", _i32), ("_pad", _i32)]


class BlurJob(C.Structure):
    _fields_ = [("src", _vp), ("dst", _vp), ("tmp", _vp), ("taps", _vp * 3), ("ntaps", _i32 * 3), ("_pad", _i32)]


class ResampleJob(C.Structure):
    _fields_ = [("src", _vp), ("dst", _vp), ("tab", _vp * 3), ("noise", _vp), ("rng", Rng), ("noise_std", _f32), ("has_noise", _i32),
                ("n", _i32 * 3), ("_pad", _i32)]


class NoiseJob(C.Structure):
    _fields_ = [("src", _vp), ("dst", _vp), ("noise", _vp), ("rng", Rng), ("noise_std", _f32), ("flags", _i32)]


class ZoomJob(C.Structure):
    _fields_ = [("src", _vp), ("dst", _vp), ("tab", _vp * 3), ("minmax", _vp), ("n", _i32 * 3), ("post", _i32)]


class SepAxis(C.Structure):
    _fields_ = [("q0", _vp), ("w", _vp), ("n_out", _i32), ("width", _i32), ("pos", _vp), ("taps", _vp), ("ntaps", _i32), ("_pad", _i32)]


class SepconvJob(C.Structure):
    _fields_ = [("src", _vp), ("dst", _vp), ("tmp1", _vp), ("tmp2", _vp), ("ax", SepAxis * 3), ("noise", _vp), ("rng", Rng),
                ("noise_std", _f32), ("has_noise", _i32), ("cap_dst", _i64), ("cap_tmp1", _i64), ("cap_tmp2", _i64)]


class SampleJob(C.Structure):
    _fields_ = [("labels", _vp), ("labels2", _vp), ("centers_out", _vp), ("count_out", _vp), ("workspace", _vp), ("workspace_bytes", _i64), ("rng", Rng),
                ("prior_centers", (_f32 * 3) * 4), ("prior_sigmas", (_f32 * 3) * 4), ("match", _i32), ("nprior", _i32), ("k", _i32), ("_pad", _i32)]


class PerlinOctave(C.Structure):
    _fields_ = [("grad", _vp), ("lin", _vp * 3), ("res", _i32 * 3), ("amp", _f32)]


class GridJob(C.Structure):
    _fields_ = [("out", _vp), ("rng", Rng), ("scale", _f32), ("n", _i32)]


class EmJob(C.Structure):
    _fields_ = [("x", _vp), ("index", _vp), ("n", _i64), ("k", _i32), ("label_base", _i32), ("params", _vp), ("trace", _vp), ("state", _vp), ("seeds", _vp),
                ("labels", _vp), ("out", _vp), ("rng_seed", C.c_uint64), ("rng_stream", C.c_uint64)]


class UnpackJob(C.Structure):
    _fields_ = [("words", _vp), ("out", _vp), ("shift", _i32 * 4), ("mask", _i32 * 4), ("word_bytes", _i32), ("_pad", _i32)]


class SepComposeJob(C.Structure):
    _fields_ = [("pos", _vp), ("taps", _vp), ("q0_out", _vp), ("w_out", _vp), ("ntaps", _i32), ("n_in", _i32), ("n_out", _i32), ("width", _i32), ("cap_q0", _i32), ("cap_w", _i32)]


class StepJobs(C.Structure):  # struct fsg_step_jobs
    _fields_ = [("n_gmm", _i32 * 2), ("n_grid", _i32), ("n_warp", _i32), ("n_shift", _i32), ("n_sep", _i32), ("n_noise", _i32), ("n_scale", _i32),
                ("ring_used", _i32), ("_pad", _i32), ("gmm", (GmmJob * MAX_JOBS) * 2), ("grid", GridJob * (2 * MAX_JOBS)), ("warp", WarpJob * MAX_JOBS),
                ("shift", WarpJob * MAX_JOBS), ("compose", SepComposeJob * (3 * MAX_JOBS)), ("sep", SepconvJob * MAX_JOBS), ("zoom", ZoomJob * MAX_JOBS),
                ("noise", NoiseJob * MAX_JOBS), ("scale_idx", _i32 * MAX_JOBS)]


_NP_SCALARS = {_vp: "u8", _f32: "f4", _i32: "i4", _i64: "i8", C.c_uint32: "u4", C.c_uint64: "u8", C.c_double: "f8", C.c_int16: "i2", C.c_uint8: "u1"}
_np_dtypes: dict = {}


def np_dtype(ct):
    """numpy dtype with the exact layout of a ctypes job struct (explicit offsets and item size), so that an array
    of jobs can be filled column by column and handed to the library as is."""
    import numpy as np

    dt = _np_dtypes.get(ct)
    if dt is None:
        if isinstance(ct, type) and issubclass(ct, C.Structure):
            names = [n for n, _ in ct._fields_]
            dt = np.dtype({"names": names, "formats": [np_dtype(t) for _, t in ct._fields_], "offsets": [getattr(ct, n).offset for n in names], "itemsize": C.sizeof(ct)})
        elif isinstance(ct, type) and issubclass(ct, C.Array):
            dt = np.dtype((np_dtype(ct._type_), (ct._length_,)))
        else:
            dt = np.dtype(_NP_SCALARS[ct])
        _np_dtypes[ct] = dt
    return dt


_STRUCTS = {"fsg_em_job": EmJob, "fsg_unpack_job": UnpackJob, "fsg_grid_job": GridJob, "fsg_sample_job": SampleJob, "fsg_perlin_octave": PerlinOctave, "fsg_sepaxis": SepAxis, "fsg_sepconv_job": SepconvJob, "fsg_sepcompose_job": SepComposeJob, "fsg_tab": Tab, "fsg_rng": Rng, "fsg_gmm_job": GmmJob, "fsg_warp_job": WarpJob, "fsg_blur_job": BlurJob,
            "fsg_resample_job": ResampleJob, "fsg_noise_job": NoiseJob, "fsg_zoom_job": ZoomJob,
            "fsg_texvol": TexVol, "fsg_draw_config": DrawConfig, "fsg_draw_out": DrawOut, "fsg_step_inputs": StepInputs, "fsg_step_sample": StepSample, "fsg_step": Step, "fsg_step_jobs": StepJobs}

# name -> (restype, argtypes); every symbol include/fsg.h declares
SIGNATURES = {
    "fsg_version": (C.c_int, []),
    "fsg_last_error": (C.c_char_p, []),
    "fsg_sizeof": (C.c_int, [C.c_char_p]),
    "fsg_gmm": (C.c_int, [C.POINTER(GmmJob), C.c_int, _i64, _vp]),
    "fsg_draw_batch": (C.c_int, [C.POINTER(DrawConfig), _vp, C.c_int, C.c_uint64, C.POINTER(DrawOut)]),
    "fsg_gaussian_taps": (C.c_int, [C.c_double, _vp, C.c_int]),
    "fsg_step_fill": (C.c_int, [C.POINTER(Step), C.POINTER(DrawConfig), C.POINTER(DrawOut), _vp, C.POINTER(StepInputs), C.POINTER(StepSample), _vp]),
    "fsg_step_build": (C.c_int, [C.POINTER(Step), C.POINTER(StepSample), C.POINTER(StepJobs)]),
    "fsg_step_run": (C.c_int, [C.POINTER(Step), C.POINTER(StepSample), _vp]),
    "fsg_texvol_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(TexVol)]),
    "fsg_texvol_destroy": (C.c_int, [C.POINTER(TexVol)]),
    "fsg_texvol_copy": (C.c_int, [C.POINTER(TexVol), _vp, C.c_int, _vp]),
    "fsg_warp_shift": (C.c_int, [C.POINTER(WarpJob), C.c_int, C.c_int, C.c_int, C.c_int, _vp]),
    "fsg_warp": (C.c_int, [C.POINTER(WarpJob), C.c_int, C.c_int, C.c_int, C.c_int, _vp]),
    "fsg_warp_coords": (C.c_int, [C.POINTER(WarpJob), C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
    "fsg_blur3d": (C.c_int, [C.POINTER(BlurJob), C.c_int, C.c_int, C.c_int, C.c_int, _vp]),
    "fsg_sep_compose": (C.c_int, [C.POINTER(SepComposeJob), C.c_int, _vp]),
    "fsg_sepconv": (C.c_int, [C.POINTER(SepconvJob), C.c_int] + [C.c_int] * 3 + [_vp]),
    "fsg_resample": (C.c_int, [C.POINTER(ResampleJob), C.c_int] + [C.c_int] * 3 + [_vp]),
    "fsg_add_noise": (C.c_int, [C.POINTER(NoiseJob), C.c_int, _i64, _vp]),
    "fsg_zoom_minmax": (C.c_int, [C.POINTER(ZoomJob), C.c_int] + [C.c_int] * 3 + [_vp]),
    "fsg_zoom": (C.c_int, [C.POINTER(ZoomJob), C.c_int] + [C.c_int] * 3 + [_vp]),
    "fsg_minmax": (C.c_int, [_vp, _i64, _vp, _vp]),
    "fsg_scale_intensity": (C.c_int, [_vp, _vp, _i64, _vp, _vp]),
    "fsg_f32_to_u8": (C.c_int, [_vp, _vp, _i64, _vp]),
    "fsg_u8_to_f32": (C.c_int, [_vp, _vp, _i64, _vp]),
    "fsg_u8_to_i64": (C.c_int, [_vp, _vp, _i64, _vp]),
    "fsg_mog": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp]),
    "fsg_sample_voxels": (C.c_int, [C.POINTER(SampleJob), C.c_int, C.c_int, C.c_int, _vp]),
    "fsg_perlin": (C.c_int, [C.POINTER(PerlinOctave), C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp]),
    "fsg_struct_blend": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _f32, _f32, _vp, _i64, _vp]),
    "fsg_morph_box": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp]),
    "fsg_morph_dist": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp]),
    "fsg_morph_thresh": (C.c_int, [_vp, _vp, C.c_int, _i64, _vp]),
    "fsg_morph_ring": (C.c_int, [_vp, _vp, _vp, Rng, _f32, _vp, _i64, _vp]),
    "fsg_morph_count_merge": (C.c_int, [_vp, _vp, C.c_int, _vp, _i64, _vp]),
    "fsg_boundary_select": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_int, _vp, _vp, _i64, _vp]),
    "fsg_mask_mul": (C.c_int, [_vp, _vp, _vp, _i64, _vp]),
    "fsg_label_mask": (C.c_int, [_vp, C.c_int, _vp, _i64, _vp]),
    "fsg_slice_acq_forward": (C.c_int, [_vp, _vp, _vp, C.c_int, _f32, _vp] + [C.c_int] * 6 + [_f32, _vp]),
    "fsg_slice_acq_forward_xpairs": (C.c_int, [_vp, _vp, _vp, C.c_int, _f32, _vp] + [C.c_int] * 6 + [_f32, _vp]),
    "fsg_slice_acq_forward_xyquads": (C.c_int, [_vp, _vp, _vp, C.c_int, _f32, _vp] + [C.c_int] * 6 + [_f32, _vp]),
    "fsg_volume_xyquads": (C.c_int, [_vp, _vp, _i64, C.c_int, _vp]),
    "fsg_volume_xpairs": (C.c_int, [_vp, _vp, _i64, _vp]),
    "fsg_slice_acq_adjoint": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, _vp, C.c_int, _f32, _vp, _vp, _vp, _vp, _vp] + [C.c_int] * 6 + [_f32, C.c_int, _vp]),
    "fsg_slice_acq_forward_ex": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp, C.c_int, _f32, _vp, _vp] + [C.c_int] * 6 + [_f32, C.c_int, _vp]),
    "fsg_slice_acq_adjoint_ex": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, _vp, C.c_int, _f32, _vp, _vp, _vp, _vp, _vp] + [C.c_int] * 6 + [_f32, C.c_int, C.c_int, _vp]),
    "fsg_axisangle2mat": (C.c_int, [_vp, _vp, C.c_int, _vp]),
    "fsg_mat2axisangle": (C.c_int, [_vp, _vp, C.c_int, _vp]),
    "fsg_slice_sums": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp]),
    "fsg_slice_gamma": (C.c_int, [_vp, _i64, _f32, _vp, _vp]),
    "fsg_slice_rician": (C.c_int, [_vp, _i64, _f32, _f32, _vp, _vp, Rng, _vp]),
    "fsg_slice_void": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp, C.c_int, _vp]),
    "fsg_recon_merge": (C.c_int, [_vp, _vp, _vp, _vp, _f32, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "fsg_philox_fill": (C.c_int, [Rng, _vp, _i64, C.c_int, _vp]),
    "fsg_draw_grids": (C.c_int, [C.POINTER(GridJob), C.c_int, _vp]),
    "fsg_fetch_params": (C.c_int, [_vp, _vp, _i64, _vp]),
    "fsg_unpack_seeds": (C.c_int, [C.POINTER(UnpackJob), C.c_int, _i64, _vp]),
    "fsg_seed_partition_workspace": (_i64, [_i64]),
    "fsg_seed_partition": (C.c_int, [_vp, _vp, C.c_char_p, _i64, _vp, _vp, _vp, _vp, _i64, _vp]),
    "fsg_em_workspace": (_i64, [C.c_int]),
    "fsg_em_seed": (C.c_int, [C.POINTER(EmJob), C.c_int, _vp, _i64, _vp]),
    "fsg_em_fit": (C.c_int, [C.POINTER(EmJob), C.c_int, C.c_int, C.c_double, C.c_double, _vp, _i64, _vp]),
    "fsg_em_predict": (C.c_int, [C.POINTER(EmJob), C.c_int, _vp, _i64, _vp]),
}

# kernels one call of an entry point launches on the fused base path (profiles/r01g_launches.csv);
# entry points not listed launch one
KERNELS_PER_CALL = {"fsg_warp_shift": 4, "fsg_sepconv": 3, "fsg_zoom_minmax": 3, "fsg_minmax": 3, "fsg_slice_acq_adjoint": 2, "fsg_slice_gamma": 2,
                    "fsg_texvol_create": 0, "fsg_texvol_destroy": 0, "fsg_texvol_copy": 0,
                    "fsg_step_build": 0, "fsg_step_run": 16, "fsg_draw_batch": 0, "fsg_gaussian_taps": 0, "fsg_step_fill": 0}

_lib = None


class FsgError(RuntimeError):
    pass


def load(build_if_missing: bool = True):
    """Load libfsg.so (building it with nvcc when absent and a compiler is available)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        if not build_if_missing:
            raise FsgError(f"{LIB_PATH} is missing: run `python -m fetalsyngen_b200.build` (there is no CPU fallback)")
        from . import build as _b

        _b.build()
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype, fn.argtypes = res, args
    for cname, st in _STRUCTS.items():
        got = lib.fsg_sizeof(cname.encode())
        if got != C.sizeof(st):
            raise FsgError(f"ABI mismatch for {cname}: library {got} bytes, binding {C.sizeof(st)} bytes")
    _lib = lib
    return lib


class LaunchStats:
    """Counts C-ABI calls and, when ``timing`` is on, brackets each call with CUDA events on the
    launching stream (bench.py uses this for the per-kernel roofline)."""

    def __init__(self):
        self.calls: dict = {}
        self.timing = False     # True: every entry point; a set of names: only those
        self._events: list = []

    def reset(self):
        self.calls.clear()
        self._events.clear()

    def total_calls(self) -> int:
        return sum(self.calls.values())

    def total_kernels(self) -> int:
        """Kernel launches behind the counted calls (memsets and copies not counted)."""
        return sum(n * KERNELS_PER_CALL.get(name, 1) for name, n in self.calls.items())

    def elapsed_ms(self) -> dict:
        """Per entry point: (number of calls, total device milliseconds). Synchronises."""
        import torch

        torch.cuda.synchronize()
        out: dict = {}
        for name, a, b in self._events:
            n, t = out.get(name, (0, 0.0))
            out[name] = (n + 1, t + a.elapsed_time(b))
        return out


stats = LaunchStats()


def call(name: str, *args):
    lib = load()
    stats.calls[name] = stats.calls.get(name, 0) + 1
    if stats.timing is True or (stats.timing and name in stats.timing):
        import torch

        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        rc = getattr(lib, name)(*args)
        b.record()
        stats._events.append((name, a, b))
    else:
        rc = getattr(lib, name)(*args)
    if rc != 0:
        raise FsgError(f"{name} failed ({rc}): {lib.fsg_last_error().decode()}")
