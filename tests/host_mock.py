"""CPU stand-in for the device engine, for tests and profiling of the HOST side only: the engine's scratch volumes
are CPU tensors and every C-ABI call is intercepted (recorded or dropped), so the Python that draws parameters,
fills job structs and queues launches runs without a GPU.  Nothing here computes a volume."""
import ctypes as C

import numpy as np
import torch

from fetalsyngen_b200 import _lib, engine as E
from fetalsyngen_b200.data.packed import PackedSeeds
from fetalsyngen_b200.tables import DeviceTables


class _Event:
    def __init__(self, *a, **k):
        pass

    def record(self, *a):
        pass

    def synchronize(self):
        pass


class Recorder:
    """Replaces ``_lib.call``: keeps (name, bytes of the job array, scalar arguments) of every call."""

    def __init__(self, keep=True):
        self.keep, self.calls = keep, []

    def __call__(self, name, *args):
        if not self.keep:
            return
        if name == "fsg_texvol_create":
            h = args[3]._obj
            n = len(self.calls) + 1
            h.array, h.tex, h.surf = 7000 + n, 8000 + n, 9000 + n
            return
        if name in ("fsg_texvol_destroy", "fsg_fetch_params"):
            return
        struct = {"fsg_gmm": _lib.GmmJob, "fsg_draw_grids": _lib.GridJob, "fsg_warp_shift": _lib.WarpJob, "fsg_warp": _lib.WarpJob, "fsg_sep_compose": _lib.SepComposeJob,
                  "fsg_sepconv": _lib.SepconvJob, "fsg_zoom_minmax": _lib.ZoomJob, "fsg_zoom": _lib.ZoomJob, "fsg_add_noise": _lib.NoiseJob}.get(name)
        if struct is None:
            self.calls.append((name, None, tuple(a for a in args[:-1])))
            return
        n = int(args[1])
        raw = C.string_at(C.cast(args[0], C.c_void_p).value, n * C.sizeof(struct))
        self.calls.append((name, np.frombuffer(raw, dtype=_lib.np_dtype(struct)).copy(), tuple(int(a) for a in args[2:-1])))


def install(recorder):
    """Route the engine's C-ABI calls to `recorder` (a callable); returns a function that restores the originals."""
    saved = (_lib.call, E._stream, torch.cuda.Event)
    _lib.load()
    _lib.call = recorder
    E._stream = lambda: 0
    torch.cuda.Event = _Event

    def restore():
        _lib.call, E._stream, torch.cuda.Event = saved

    return restore


def cpu_engine(shape, resolution):
    eng = object.__new__(E.SynthEngine)
    eng.shape = tuple(int(s) for s in shape)
    eng.resolution = np.asarray(resolution, dtype=np.float64)
    eng.device = torch.device("cpu")
    eng.nvox = int(np.prod(eng.shape))
    eng.tables = DeviceTables(eng.device)
    eng._scratch, eng._batch, eng.use_tex, eng._texvols, eng._ptrs = {}, None, True, [], {}
    hring = torch.empty((eng.RING_SLOTS, eng.RING_FLOATS), dtype=torch.float32)
    eng._ring = (hring, torch.empty_like(hring), [None] * eng.RING_SLOTS, [0])
    eng._ring_np = hring.numpy()
    return eng


class FakePacked(PackedSeeds):
    """A packed subject without data: layout and a word buffer of the right size."""

    def __init__(self, shape):
        self.shape, self.word_bytes, self.counts = tuple(shape), 2, list(range(1, 7))
        self.layout = {n: (3 + 2 * (n - 1), 7) for n in self.counts}
        self._w = torch.zeros(int(np.prod(shape)), dtype=torch.int16)

    def on(self, device):
        return self._w
