"""CPU tests (no GPU): the C-ABI library loads and exports what include/fsg.h declares, the
product has no CPU fallback and never touches oracle/, host-side logic (config instantiation,
NIfTI codec, dataset discovery, sample sharding) and the multi-process (gloo, world_size 2)
layout of the benchmark."""
import os
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
REF_CFG = Path("/root/reference/configs")


# ----------------------------------------------------------------------------- boundary
def _declared():
    text = (ROOT / "include" / "fsg.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fsg_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from fetalsyngen_b200 import _lib

    names = _declared()
    assert len(names) >= 30
    assert set(names) == set(_lib.SIGNATURES), (set(names) ^ set(_lib.SIGNATURES))
    lib = _lib.load(build_if_missing=False)  # also checks struct sizes against fsg_sizeof
    for n in names:
        assert hasattr(lib, n)
    out = subprocess.run(["nm", "-D", "--defined-only", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (fsg_[a-z0-9_]+)", out))
    assert set(names) <= exported
    assert lib.fsg_version() >= 100
    assert isinstance(lib.fsg_last_error(), bytes)


def test_integration_doc_names_every_entry_point():
    doc = (ROOT / "INTEGRATION.md").read_text()
    assert [n for n in _declared() if n not in doc] == []


def test_no_cpu_fallback_and_no_oracle_in_product():
    from fetalsyngen_b200 import _lib
    from fetalsyngen_b200.engine import SynthEngine

    with pytest.raises(_lib.FsgError):
        SynthEngine((8, 8, 8), (1.0, 1.0, 1.0), "cpu")
    if not torch.cuda.is_available():
        with pytest.raises(_lib.FsgError):
            SynthEngine((8, 8, 8), (1.0, 1.0, 1.0), "cuda:0")
    for p in (ROOT / "fetalsyngen_b200").rglob("*.py"):
        src = p.read_text()
        assert "np_oracle" not in src and "np_motion" not in src and "np_artifacts" not in src and "np_seeds" not in src and "ref_import" not in src, p
        assert "import sklearn" not in src and "from sklearn" not in src, p  # the reference's clustering dependency is a checker, never a code path


def test_only_tests_smoke_and_bench_touch_the_oracle():
    """oracle/ is test infrastructure: tools/ and the package never import it (bench.py and
    __graft_entry__.py hold the only measurement / smoke legs that do)."""
    names = ("np_oracle", "np_motion", "np_artifacts", "np_seeds", "ref_import", "build_ref")
    for folder in ("tools", "fetalsyngen_b200"):
        for p in (ROOT / folder).rglob("*.py"):
            src = p.read_text()
            for n in names:
                assert f"import {n}" not in src, (p, n)


def test_c_abi_argument_checks_without_a_gpu():
    """Argument validation happens before any CUDA call: bad arguments give rc != 0 + a message."""
    from fetalsyngen_b200 import _lib

    lib = _lib.load(build_if_missing=False)
    assert lib.fsg_slice_sums(None, 0, 0, None, None) != 0
    assert b"fsg_slice_sums" in lib.fsg_last_error()
    assert lib.fsg_gmm(None, 0, 0, None) != 0
    assert lib.fsg_sizeof(b"no_such_struct") <= 0


# ----------------------------------------------------------------------------- config
@pytest.mark.skipif(not REF_CFG.exists(), reason="reference configs not present")
def test_reference_yaml_builds_our_classes():
    from fetalsyngen_b200 import config
    from fetalsyngen_b200.generator.augmentation.artifacts import BlurCortex, SimulatedBoundaries, SimulateMotion, StructNoise
    from fetalsyngen_b200.generator.model import FetalSynthGen

    cfg = config.load_yaml(REF_CFG / "dataset/generator/default.yaml")
    cfg["device"] = "cuda:0"
    gen = config.instantiate(cfg)
    assert isinstance(gen, FetalSynthGen)
    assert list(gen.shape) == [256, 256, 256] and gen.spatial_deform.device == "cuda:0"
    assert isinstance(gen.artifacts["blur_cortex"], BlurCortex) and isinstance(gen.artifacts["struct_noise"], StructNoise)
    assert isinstance(gen.artifacts["simulate_motion"], SimulateMotion) and isinstance(gen.artifacts["boundaries"], SimulatedBoundaries)
    assert gen.artifacts["simulate_motion"].recon_args.merge_params.merge_type == "perlin"
    assert gen.artifacts["simulate_motion"].scanner_args.max_num_slices == 250


def test_instantiate_interpolation_and_errors(tmp_path):
    from fetalsyngen_b200 import config

    cfg = {"device": "cuda:1", "inner": {"_target_": "fetalsyngen.generator.augmentation.synthseg.RandGamma", "prob": 0.5, "gamma_std": 0.1}, "d2": "${.device}",
           "nest": {"dev": "${..device}"}}
    out = config.instantiate(cfg)
    assert out["d2"] == "cuda:1" and out["nest"]["dev"] == "cuda:1"
    assert type(out["inner"]).__module__.startswith("fetalsyngen_b200.")
    with pytest.raises(ImportError):
        config.instantiate({"_target_": "fetalsyngen.generator.model.NoSuchClass"})
    (tmp_path / "generator").mkdir()
    (tmp_path / "generator/default.yaml").write_text("shape: [4, 4, 4]\ndevice: cpu\n")
    (tmp_path / "ds.yaml").write_text("defaults:\n  - generator/default\nname: x\ngenerator:\n  device: cuda\n")
    merged = config.load_yaml(tmp_path / "ds.yaml")
    assert merged["generator"] == {"shape": [4, 4, 4], "device": "cuda"} and merged["name"] == "x"


# ----------------------------------------------------------------------------- NIfTI + dataset
def test_nifti_round_trip_and_errors(tmp_path):
    from fetalsyngen_b200.utils import nifti

    rs = np.random.RandomState(0)
    for dt in (np.float32, np.int8, np.uint8, np.int16):
        a = (rs.rand(5, 6, 7) * 100).astype(dt)
        aff = np.diag([0.5, 0.5, 0.5, 1.0])
        nifti.write_nifti(tmp_path / "a.nii.gz", a, aff)
        b, aff2 = nifti.read_nifti(tmp_path / "a.nii.gz", with_affine=True)
        assert b.dtype == dt and np.array_equal(a, b) and np.allclose(aff, aff2)
    (tmp_path / "bad.nii").write_bytes(b"\x00" * 100)
    with pytest.raises(nifti.NiftiError):
        nifti.read_nifti(tmp_path / "bad.nii")


def _bids(tmp_path, subs=("sub-a", "sub-b"), with_img=True):
    from fetalsyngen_b200.utils import nifti

    for s in subs:
        d = tmp_path / "bids" / s / "anat"
        d.mkdir(parents=True)
        seg = np.zeros((8, 8, 8), np.float32)
        seg[2:6, 2:6, 2:6] = 3
        nifti.write_nifti(d / f"{s}_rec-x_T2w_dseg.nii.gz", seg, np.diag([0.5, 0.5, 0.5, 1]))
        if with_img:
            nifti.write_nifti(d / f"{s}_rec-x_T2w.nii.gz", seg * 10, np.diag([0.5, 0.5, 0.5, 1]))
        for n in range(1, 3):
            for m in range(1, 5):
                sd = tmp_path / "seeds" / f"subclasses_{n}" / s / "anat"
                sd.mkdir(parents=True, exist_ok=True)
                v = np.zeros((8, 8, 8), np.int8)
                v[m : m + 1] = 10 * m
                nifti.write_nifti(sd / f"{s}_rec-x_T2w_dseg_mlabel_{m}.nii.gz", v, np.diag([0.5, 0.5, 0.5, 1]))
    return tmp_path / "bids", tmp_path / "seeds"


def test_dataset_discovery_and_errors(tmp_path):
    from fetalsyngen_b200.data.datasets import FetalSynthDataset, FetalTestDataset

    bids, seeds = _bids(tmp_path)
    ds = FetalTestDataset(str(bids), None)
    assert len(ds) == 2 and ds.subjects == ["sub-a", "sub-b"]
    ds1 = FetalTestDataset(str(bids), ["sub-b", "sub-zzz"])
    assert ds1.subjects == ["sub-b"]
    item = ds[0]
    assert item["image"].shape == (1, 8, 8, 8) and item["label"].shape == (1, 8, 8, 8) and item["name"] == "sub-a"
    assert float(item["image"].max()) == 30.0 and item["label"].dtype == torch.int64  # no transforms: raw intensities (datasets.py:150-176)

    class G:  # the dataset only stores the generator at construction
        shape, resolution, device = [8, 8, 8], [0.5] * 3, "cuda:0"

    sd = FetalSynthDataset(str(bids), G(), str(seeds), None, load_image=False)
    assert len(sd) == 2
    assert sorted(sd.seed_paths["sub-a"].keys()) == [1, 2] and sorted(sd.seed_paths["sub-a"][1].keys()) == [1, 2, 3, 4]
    with pytest.raises(Exception):
        FetalSynthDataset(str(bids), G(), str(tmp_path / "nope"), None)


# ----------------------------------------------------------------------------- sharding
def test_shard_ids_cover_and_are_disjoint():
    from fetalsyngen_b200 import sharding as S

    for world in (1, 2, 3, 8):
        got = sorted(i for r in range(world) for i in S.shard_ids(5, 64, r, world))
        assert got == list(range(5, 69))
        step = sorted(i for r in range(world) for i in S.step_ids(3, 8, r, world))
        assert step == list(range(3 * 8 * world, 4 * 8 * world))
    with pytest.raises(ValueError):
        S.shard_ids(0, 4, 2, 2)
    seeds = {S.sample_seed(1234, i) for i in range(10000)}
    assert len(seeds) > 9990 and all(0 <= s < 2**32 for s in seeds)
    assert S.sample_seed(1, 2) != S.sample_seed(2, 1)


_WORKER = r'''
import os, sys, json
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
from fetalsyngen_b200 import sharding as S
sys.path.insert(0, os.path.join(sys.argv[1]))
import bench
dist.init_process_group("gloo")
rank, world, local = S.rank_info()
ids = [i for st in range(2) for i in S.step_ids(st, 4, rank, world)]
sig = {}
for i in ids:                      # the host draws of a sample depend on (base_seed, id) only
    np.random.seed(S.sample_seed(1234, i)); torch.default_generator.manual_seed(S.sample_seed(1234, i))
    q = bench.draw_oracle_params(np.random, (16, 16, 16))
    sig[i] = float(q["A"].sum() + q["mus"].sum() + q["noise_std"])
allsig = [None] * world
dist.all_gather_object(allsig, sig)
S.barrier()
t = S.max_over_ranks(10.0 + rank)  # slowest rank defines the step time
if rank == 0:
    merged = {}
    for d in allsig: merged.update(d)
    print(json.dumps({"n": len(merged), "ids": sorted(merged), "max": t, "sig": [merged[k] for k in sorted(merged)]}))
dist.destroy_process_group()
'''


@pytest.mark.timeout(300)
def test_two_process_gloo_layout(tmp_path):
    """world_size 2 over gloo: ranks own disjoint ids that cover the job, per-sample draws equal the
    single-process ones (independent of the world size), timing is the max over ranks."""
    import json

    from fetalsyngen_b200 import sharding as S

    sys.path.insert(0, str(ROOT))
    import bench

    w = tmp_path / "worker.py"
    w.write_text(_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29531", str(w), str(ROOT)]
    res = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=280)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads([l for l in res.stdout.splitlines() if l.startswith("{")][-1])
    assert line["ids"] == list(range(16)) and line["max"] == 11.0
    want = []
    for i in range(16):
        np.random.seed(S.sample_seed(1234, i))
        q = bench.draw_oracle_params(np.random, (16, 16, 16))
        want.append(float(q["A"].sum() + q["mus"].sum() + q["noise_std"]))
    assert np.allclose(line["sig"], want, rtol=0, atol=0)


def test_bench_reference_arm_nonzero_rank_exits_without_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, env=env, timeout=120)
    assert res.returncode == 0 and res.stdout.strip() == ""


# ----------------------------------------------------------------------------- vectorised batch draws
def test_batch_draw_matches_the_stage_classes_statistically():
    """batch_draw.draw_plans (one counter-based draw for a whole step) must follow the same
    distributions as the per-sample stage draws that mirror the reference (KS tests on every scalar,
    moments of the GMM tables) and be a pure function of (base_seed, sample id)."""
    from scipy.stats import ks_2samp

    sys.path.insert(0, str(ROOT))
    import bench
    from fetalsyngen_b200.batch_draw import draw_plans, uniforms
    from fetalsyngen_b200.engine import SamplePlan

    shape = (256, 256, 256)
    gen = bench.build_generator(shape, "cuda:0")
    for st in (gen.spatial_deform, gen.resampled, gen.biasfield, gen.noise, gen.gamma):
        st.prob = 0.8
    n = 3000
    plans, params = draw_plans(gen, list(range(n)), 99, shape)
    # determinism / independence of the batch composition
    p2, _ = draw_plans(gen, [7, 2999, 0], 99, shape)
    for a, b in zip(p2, (plans[7], plans[2999], plans[0])):
        assert np.array_equal(a.mus, b.mus) and (a.A is None) == (b.A is None) and (a.A is None or np.array_equal(a.A, b.A)) and a.gamma == b.gamma
    assert not np.array_equal(draw_plans(gen, [7], 100, shape)[0][0].mus, plans[7].mus)
    u = uniforms(1, np.arange(2000), 50)
    assert 0 < u.min() and u.max() < 1 and abs(u.mean() - 0.5) < 0.01 and abs(np.corrcoef(u[:, 0], u[:, 1])[0, 1]) < 0.08

    np.random.seed(5)
    torch.manual_seed(5)
    ref = []
    for _ in range(n):
        p = SamplePlan()
        gen._draw_generate(p, None, shape, {}, None, device_grids=True)
        p.mus, p.sigmas = gen.intensity_generator.draw_gmm({})
        gen._draw_augment(p, shape, {}, None, device_grids=True)
        ref.append(p)

    def col(ps, f):
        return np.array([f(p) for p in ps if f(p) is not None], dtype=np.float64)

    feats = {
        "A00": lambda p: None if p.A is None else p.A[0, 0], "A01": lambda p: None if p.A is None else p.A[0, 1], "A21": lambda p: None if p.A is None else p.A[2, 1],
        "A22": lambda p: None if p.A is None else p.A[2, 2], "nonlin_std": lambda p: None if p.fsmall_dev is None else p.fsmall_dev[1],
        "fsize": lambda p: None if p.fsmall_dev is None else p.fsmall_dev[0][0], "gamma": lambda p: p.gamma,
        "bf_std": lambda p: None if p.bf_dev is None else p.bf_dev[1], "bf_size": lambda p: None if p.bf_dev is None else p.bf_dev[0][1],
        "spacing": lambda p: None if p.spacing is None else p.spacing[0], "std0": lambda p: None if p.stds is None else p.stds[0], "noise_std": lambda p: p.noise_std,
        "mus0": lambda p: p.mus[0], "mus15": lambda p: p.mus[15], "mus45": lambda p: p.mus[45], "sig22": lambda p: p.sigmas[22],
    }
    for name, f in feats.items():
        a, b = col(plans, f), col(ref, f)
        assert abs(len(a) - len(b)) < 5 * np.sqrt(n * 0.8 * 0.2) + 1, (name, len(a), len(b))  # gate probabilities
        assert ks_2samp(a, b).pvalue > 1e-4, (name, ks_2samp(a, b))
    flips = np.mean([p.flip for p in plans if p.deform]), np.mean([p.flip for p in ref if p.deform])
    assert abs(flips[0] - flips[1]) < 0.06
    # c2 = centre of the volume when shape == size; float64
    assert all(np.allclose(p.c2, 127.5) and p.c2.dtype == np.float64 for p in plans if p.deform)
    # parameter dictionaries keep the reference's keys
    assert set(params[0]) == {"selected_seeds", "seed_intensities", "deform_params", "gamma_params", "bf_params", "resample_params", "noise_params"}


def test_pack_dataset_tool_writes_one_cache_file_per_subject(tmp_path):
    """tools/pack_dataset.py: the one-time converter of a BIDS + seeds tree into the bit-packed cache."""
    import importlib.util

    from fetalsyngen_b200.data.packed import load_packed, unpack_numpy

    bids, seeds = _bids(tmp_path)
    spec = importlib.util.spec_from_file_location("pack_dataset", ROOT / "tools" / "pack_dataset.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.main(["--bids_path", str(bids), "--seed_path", str(seeds), "--out_path", str(tmp_path / "cache")]) == 0
    for s in ("sub-a", "sub-b"):
        seg, ps, _ = load_packed(tmp_path / "cache" / f"{s}.fsgpack.npz")
        assert seg.shape == (8, 8, 8) and seg.max() == 3 and ps.counts == [1, 2]
        lab = unpack_numpy(ps._host, ps.counts, {1: 2, 2: 1, 3: 2, 4: 1})
        for m in range(1, 5):
            assert (lab[m : m + 1] == 10 * m).all()


def test_volumes_are_reoriented_to_ras(tmp_path):
    """The reference applies monai Orientation('RAS') to every volume it loads (datasets.py:284-286,
    rand_gmm.py:91-96): LPS / permuted exports must come back in the same storage order as their RAS twin."""
    from fetalsyngen_b200.utils.image_reading import SimpleITKReader
    from fetalsyngen_b200.utils.nifti import ras_axes, to_ras, write_nifti

    rs = np.random.RandomState(0)
    ras = rs.randint(0, 8, size=(5, 6, 7)).astype(np.float32)
    aff = np.diag([0.5, 0.5, 0.5, 1.0])
    aff[:3, 3] = (-10, -20, -30)
    assert ras_axes(aff) == ((0, 1, 2), (False, False, False))
    assert to_ras(ras, aff)[0] is ras  # stored RAS already: no copy
    # the same volume exported LPS (x and y reversed) ...
    lps, aff_lps = ras[::-1, ::-1, :].copy(), aff.copy()
    aff_lps[:3, 0], aff_lps[:3, 1] = -aff[:3, 0], -aff[:3, 1]
    aff_lps[:3, 3] = aff[:3, 3] + aff[:3, 0] * 4 + aff[:3, 1] * 5
    # ... and with its axes stored as (S, R, A)
    sra, aff_sra = np.ascontiguousarray(ras.transpose(2, 0, 1)), aff.copy()
    aff_sra[:3, :3] = aff[:3, :3][:, [2, 0, 1]]
    reader = SimpleITKReader()
    for name, vol, a in (("lps", lps, aff_lps), ("sra", sra, aff_sra)):
        f = tmp_path / f"{name}.nii.gz"
        write_nifti(f, vol, a)
        got = reader(f)
        assert tuple(got.shape) == ras.shape and np.array_equal(got.numpy(), ras), name
        np.testing.assert_allclose(got.affine.numpy(), aff, atol=1e-5)


def test_byte_lru_evicts_least_recently_used():
    from fetalsyngen_b200.utils.lru import ByteLRU

    c = ByteLRU(100)
    c.put("a", 1, 40)
    c.put("b", 2, 40)
    assert c.get("a") == 1  # a is now the most recently used
    c.put("c", 3, 40)       # 120 bytes > 100: b goes
    assert "b" not in c and c.get("a") == 1 and c.get("c") == 3 and c.bytes == 80
    c.put("huge", 4, 1000)  # the newest entry always stays
    assert c.get("huge") == 4 and len(c) == 1


def test_packed_only_dataset_discovers_subjects(tmp_path):
    """FetalSynthDataset.from_packed: subject files alone, no BIDS tree (names with and without a session)."""
    from fetalsyngen_b200.data.datasets import FetalSynthDataset
    from fetalsyngen_b200.data.packed import save_packed

    seg = np.zeros((4, 4, 4), np.uint8)
    for name in ("sub-b_ses-02", "sub-a", "sub-b_ses-01"):
        save_packed(tmp_path / f"{name}.fsgpack.npz", seg, np.zeros((4, 4, 4), np.uint16), [1, 2, 3, 4, 5, 6])
    ds = FetalSynthDataset.from_packed(tmp_path, generator=None)
    assert ds.sub_ses == [("sub-a", None), ("sub-b", "ses-01"), ("sub-b", "ses-02")] and len(ds) == 2
    assert [ds._sub_ses_idx(i) for i in range(3)] == ["sub-a", "sub-b_ses-01", "sub-b_ses-02"]
    assert FetalSynthDataset.from_packed(tmp_path, None, sub_list=["sub-a"]).sub_ses == [("sub-a", None)]
    with pytest.raises(FileNotFoundError):
        FetalSynthDataset.from_packed(tmp_path / "nowhere", None)


def test_device_loader_epochs_do_not_repeat_sample_ids():
    """with base_seed every epoch replayed the same ids.  Ids of (epoch, step) are disjoint."""
    from fetalsyngen_b200.sharding import step_ids

    num_batches, B, world = 5, 4, 2
    seen = set()
    for epoch in range(3):
        for step in range(num_batches):
            for rank in range(world):
                ids = step_ids(epoch * num_batches + step, B, rank, world)
                assert not (seen & set(ids))
                seen |= set(ids)
    assert seen == set(range(3 * num_batches * B * world))


@pytest.mark.parametrize("prob,packed", [(1.0, True), (0.6, True), (0.6, False)])
def test_vectorised_job_builder_matches_the_per_sample_builder(prob, packed):
    """batch_step.run_base_batch (numpy structured job arrays filled column by column) must queue the same C-ABI calls
    with byte-identical job structs as SynthEngine.run_base (one ctypes struct per sample), for every combination of
    the per-sample gates.  Runs on a CPU stand-in of the engine that records the calls (tests/host_mock.py)."""
    sys.path.insert(0, str(ROOT / "tests"))
    from fetalsyngen_b200.batch_draw import draw_batch
    from fetalsyngen_b200.batch_step import run_base_batch
    from fetalsyngen_b200.generator.augmentation.synthseg import RandBiasField, RandGamma, RandNoise, RandResample
    from fetalsyngen_b200.generator.deformation.affine_nonrigid import SpatialDeformation
    from fetalsyngen_b200.generator.intensity.rand_gmm import ImageFromSeeds
    from fetalsyngen_b200.generator.model import FetalSynthGen
    from host_mock import FakePacked, Recorder, cpu_engine, install

    shape, B = (48, 40, 56), 8
    labels = [0] + list(range(10, 50))
    classes = [0] + [10] * 10 + [20] * 10 + [30] * 10 + list(range(40, 50))
    gen = FetalSynthGen(shape=list(shape), resolution=[0.5, 0.5, 0.5], device="cpu", intensity_generator=ImageFromSeeds(1, 6, labels, classes),
                        spatial_deform=SpatialDeformation(20, 0.02, 0.1, list(shape), prob, True, 0.03, 0.06, 4, 0.5, "cpu"),
                        resampler=RandResample(prob, 0.5, 1.5), bias_field=RandBiasField(prob, 0.004, 0.02, 0.01, 0.3), noise=RandNoise(prob, 5, 15), gamma=RandGamma(prob, 0.1))
    rec = Recorder()
    restore = install(rec)
    try:
        eng = cpu_engine(shape, gen.resolution)
        nv = eng.nvox
        segs = [torch.zeros(nv, dtype=torch.uint8) for _ in range(B)]
        subj = [FakePacked(shape) for _ in range(3)]
        vols_l = [[torch.zeros(nv, dtype=torch.int8) for _ in range(1 + b % 4)] for b in range(B)]
        out_img = torch.empty((B, *shape), dtype=torch.float32)
        out_seg = torch.empty((B, *shape), dtype=torch.uint8)
        seen_gates = set()
        for step in range(12):
            ids = list(range(step * B, (step + 1) * B))
            d = draw_batch(gen, ids, 77, shape, with_subclusters=True)
            if packed:
                seeds = [(subj[b % 3], {m: int(d.m2s[b, m - 1]) for m in range(1, 5)}) for b in range(B)]
            else:
                seeds = vols_l
            for b in range(B):
                seen_gates.add((bool(d.deform_on[b]), bool(d.bias_on[b]), bool(d.res_on[b]), bool(d.noise_on[b]), bool(d.gamma_on[b])))
            scale = step % 2 == 0
            rec.calls.clear()
            eng._ring[3][0] = 0
            assert run_base_batch(eng, d, seeds, segs, out_img, out_seg, scale)
            fast = list(rec.calls)
            rec.calls.clear()
            eng._ring[3][0] = 0
            eng.run_base(d.plans(), seeds, segs, out_img=out_img, out_seg=out_seg, scale=scale)
            slow = list(rec.calls)

            def canon(calls):  # launches of one entry point may be split per kind of label source in a different order
                return sorted(calls, key=lambda c: (c[0], -1 if c[1] is None else len(c[1]), b"" if c[1] is None else c[1].tobytes()))

            assert [c[0] for c in canon(fast)] == [c[0] for c in canon(slow)], ([c[0] for c in fast], [c[0] for c in slow])
            for (name, ja, sa), (_, jb, sb) in zip(canon(fast), canon(slow)):
                if name in ("fsg_minmax", "fsg_scale_intensity"):  # the last argument is a freshly allocated min/max pair
                    sa, sb = sa[:-1], sb[:-1]
                assert sa == sb, (name, sa, sb)
                if ja is None:
                    continue
                assert ja.dtype == jb.dtype and ja.shape == jb.shape, name
                for f in ja.dtype.names:
                    if f.startswith("_pad"):
                        continue
                    assert np.array_equal(ja[f], jb[f]), (name, f, ja[f], jb[f], step)
        if prob < 1:
            assert len(seen_gates) > 8  # the draws really exercised different gate combinations
    finally:
        restore()


@pytest.mark.parametrize("prob,packed", [(1.0, True), (0.6, True), (0.6, False)])
def test_native_step_builder_matches_the_per_sample_builder(prob, packed):
    """fsg_step_build (the C++ builder behind fsg_step_run: one C-ABI call per step) must produce the job structs and
    the parameter block that SynthEngine.run_base produces, for every combination of the per-sample gates.  Host code
    only: runs without a GPU on the CPU stand-in of the engine."""
    import ctypes as C

    sys.path.insert(0, str(ROOT / "tests"))
    from fetalsyngen_b200 import _lib
    from fetalsyngen_b200.batch_draw import draw_batch
    from fetalsyngen_b200.batch_step import fill_step
    from fetalsyngen_b200.generator.augmentation.synthseg import RandBiasField, RandGamma, RandNoise, RandResample
    from fetalsyngen_b200.generator.deformation.affine_nonrigid import SpatialDeformation
    from fetalsyngen_b200.generator.intensity.rand_gmm import ImageFromSeeds
    from fetalsyngen_b200.generator.model import FetalSynthGen
    from host_mock import FakePacked, Recorder, cpu_engine, install

    shape, B = (48, 40, 56), 8
    labels = [0] + list(range(10, 50))
    classes = [0] + [10] * 10 + [20] * 10 + [30] * 10 + list(range(40, 50))
    gen = FetalSynthGen(shape=list(shape), resolution=[0.5, 0.5, 0.5], device="cpu", intensity_generator=ImageFromSeeds(1, 6, labels, classes),
                        spatial_deform=SpatialDeformation(20, 0.02, 0.1, list(shape), prob, True, 0.03, 0.06, 4, 0.5, "cpu"),
                        resampler=RandResample(prob, 0.5, 1.5), bias_field=RandBiasField(prob, 0.004, 0.02, 0.01, 0.3), noise=RandNoise(prob, 5, 15), gamma=RandGamma(prob, 0.1))
    rec = Recorder()
    restore = install(rec)
    try:
        lib = _lib.load()
        eng = cpu_engine(shape, gen.resolution)
        nv = eng.nvox
        segs = [torch.zeros(nv, dtype=torch.uint8) for _ in range(B)]
        subj = [FakePacked(shape) for _ in range(3)]
        vols_l = [[torch.zeros(nv, dtype=torch.int8) for _ in range(3)] for b in range(B)]
        out_img = torch.empty((B, *shape), dtype=torch.float32)
        out_seg = torch.empty((B, *shape), dtype=torch.uint8)
        eng.scratch("minmax", B, torch.float32, 2)  # the native builder keeps B rows (stand-alone ScaleIntensity uses the tail)
        for step in range(10):
            ids = list(range(step * B, (step + 1) * B))
            d = draw_batch(gen, ids, 91, shape, with_subclusters=True)
            seeds = [(subj[b % 3], {m: int(d.m2s[b, m - 1]) for m in range(1, 5)}) for b in range(B)] if packed else vols_l
            scale = step % 2 == 0
            # ---- generic builder (recorded)
            rec.calls.clear()
            eng._ring[3][0] = 0
            eng.run_base(d.plans(), seeds, segs, out_img=out_img, out_seg=out_seg, scale=scale)
            slow = list(rec.calls)
            # ---- native builder
            keep = []
            st, S = fill_step(eng, d, seeds, segs, out_img, out_seg, scale, keep)
            st.ring_host, st.ring_dev, st.ring_floats = eng._ring[0][0].data_ptr(), eng._ring[1][0].data_ptr(), eng.RING_FLOATS
            ring_generic = eng._ring_np[0][:4096].copy()
            J = _lib.StepJobs()
            rc = lib.fsg_step_build(C.byref(st), C.cast(S.ctypes.data, C.POINTER(_lib.StepSample)), C.byref(J))
            assert rc == 0, lib.fsg_last_error().decode()
            assert np.array_equal(eng._ring_np[0][: J.ring_used], ring_generic[: J.ring_used])  # means, sigmas, taps at the same offsets

            def arr(field, n, struct):
                return np.frombuffer(bytes(field), dtype=_lib.np_dtype(struct))[:n].copy()

            native = {
                "fsg_gmm": np.concatenate([arr(J.gmm[0], J.n_gmm[0], _lib.GmmJob), arr(J.gmm[1], J.n_gmm[1], _lib.GmmJob)]),
                "fsg_draw_grids": arr(J.grid, J.n_grid, _lib.GridJob), "fsg_warp_shift": arr(J.shift, J.n_shift, _lib.WarpJob),
                "fsg_warp": arr(J.warp, J.n_warp, _lib.WarpJob), "fsg_sep_compose": arr(J.compose, 3 * J.n_sep, _lib.SepComposeJob),
                "fsg_sepconv": arr(J.sep, J.n_sep, _lib.SepconvJob), "fsg_zoom_minmax": arr(J.zoom, J.n_sep, _lib.ZoomJob),
                "fsg_zoom": arr(J.zoom, J.n_sep, _lib.ZoomJob), "fsg_add_noise": arr(J.noise, J.n_noise, _lib.NoiseJob),
            }
            for name, want in native.items():
                got = [c[1] for c in slow if c[0] == name]
                got = np.concatenate(got) if got else want[:0]
                assert len(got) == len(want), (name, len(got), len(want), step)
                if name == "fsg_gmm":  # the generic builder launches per kind of label source; order within a kind is kept
                    got = got[np.argsort(got["mus"], kind="stable")]
                    want = want[np.argsort(want["mus"], kind="stable")]
                for f in want.dtype.names:
                    if not f.startswith("_pad"):
                        assert np.array_equal(got[f], want[f]), (name, f, got[f], want[f], step)
            scaled = [c[2][0] for c in slow if c[0] == "fsg_minmax"]
            assert scaled == [out_img.data_ptr() + 4 * nv * int(J.scale_idx[i]) for i in range(J.n_scale)]
    finally:
        restore()


def test_native_draws_match_the_numpy_definition():
    """fsg_draw_batch (the library's per-sample parameter draws) against the numpy code it was written from: gates,
    integers and everything built from +, -, *, / bit for bit; cos / sin / exp / ndtri to rounding."""
    sys.path.insert(0, str(ROOT))
    import bench
    from fetalsyngen_b200.batch_draw import draw_batch, draw_batch_numpy

    shape = (256, 256, 256)
    gen = bench.build_generator(shape, "cpu")
    ids = list(range(1000, 1256))
    a = draw_batch(gen, ids, 4242, shape, True)
    b = draw_batch_numpy(gen, ids, 4242, shape, True)
    for name in ("deform_on", "flip", "gamma_on", "bias_on", "res_on", "noise_on", "size_f", "bf_size", "m2s", "sample_ids"):
        assert np.array_equal(getattr(a, name), getattr(b, name)), name
    for name in ("mus", "sigmas", "rot", "shear", "scal", "A", "c2", "nonlin_scale", "nonlin_std", "gamma", "bf_scale", "bf_std", "spacing", "stds", "noise_std"):
        x, y = np.asarray(getattr(a, name), dtype=np.float64), np.asarray(getattr(b, name), dtype=np.float64)
        assert x.shape == y.shape and np.abs(x - y).max() <= 1e-6 * max(1.0, np.abs(y).max()), name
        assert getattr(a, name).dtype == getattr(b, name).dtype, name
    assert np.array_equal(a.spacing, b.spacing) and np.array_equal(a.stds, b.stds) and np.array_equal(a.c2, b.c2)  # pure arithmetic: exact
    # plans and parameter dictionaries come out of either
    assert len(a.plans()) == 256 and a.params()[3]["resample_params"]["spacing"] == b.params()[3]["resample_params"]["spacing"]


@pytest.mark.parametrize("prob,packed", [(1.0, True), (0.6, True), (0.6, False)])
def test_native_sample_structs_match_the_numpy_ones(prob, packed):
    """fsg_step_fill (draws + input addresses -> fsg_step_sample) against batch_step.fill_step (the same in numpy)."""
    sys.path.insert(0, str(ROOT / "tests"))
    from fetalsyngen_b200 import _lib
    from fetalsyngen_b200.batch_draw import draw_batch
    from fetalsyngen_b200.batch_step import fill_step, prepare_step_native
    from fetalsyngen_b200.generator.augmentation.synthseg import RandBiasField, RandGamma, RandNoise, RandResample
    from fetalsyngen_b200.generator.deformation.affine_nonrigid import SpatialDeformation
    from fetalsyngen_b200.generator.intensity.rand_gmm import ImageFromSeeds
    from fetalsyngen_b200.generator.model import FetalSynthGen
    from host_mock import FakePacked, Recorder, cpu_engine, install

    shape, B = (48, 40, 56), 8
    labels = [0] + list(range(10, 50))
    classes = [0] + [10] * 10 + [20] * 10 + [30] * 10 + list(range(40, 50))
    gen = FetalSynthGen(shape=list(shape), resolution=[0.5, 0.5, 0.5], device="cpu", intensity_generator=ImageFromSeeds(1, 6, labels, classes),
                        spatial_deform=SpatialDeformation(20, 0.02, 0.1, list(shape), prob, True, 0.03, 0.06, 4, 0.5, "cpu"),
                        resampler=RandResample(prob, 0.5, 1.5), bias_field=RandBiasField(prob, 0.004, 0.02, 0.01, 0.3), noise=RandNoise(prob, 5, 15), gamma=RandGamma(prob, 0.1))
    restore = install(Recorder())
    try:
        eng = cpu_engine(shape, gen.resolution)
        nv = eng.nvox
        segs = [torch.zeros(nv, dtype=torch.uint8) for _ in range(B)]
        subj = [FakePacked(shape) for _ in range(3)]
        vols_l = [[torch.zeros(nv, dtype=torch.int8) for _ in range(3)] for b in range(B)]
        out_img = torch.empty((B, *shape), dtype=torch.float32)
        out_seg = torch.empty((B, *shape), dtype=torch.uint8)
        for step in range(8):
            ids = list(range(step * B, (step + 1) * B))
            d = draw_batch(gen, ids, 17, shape, with_subclusters=True)
            raw = [subj[b % 3] for b in range(B)] if packed else vols_l
            tup = [(subj[b % 3], {m: int(d.m2s[b, m - 1]) for m in range(1, 5)}) for b in range(B)] if packed else vols_l
            keep = []
            _, S_np = fill_step(eng, d, tup, segs, out_img, out_seg, True, keep)
            prepared = prepare_step_native(eng, gen, d, raw, segs, out_img, out_seg, True)
            assert prepared is not None
            st, S_c = prepared
            got = np.frombuffer(bytes(S_c), dtype=_lib.np_dtype(_lib.StepSample))[:B]
            for f in got.dtype.names:
                if f in ("mus", "sigmas", "taps"):
                    continue  # host addresses: compared by content below
                # table fields only matter (and are only filled by the library) where the sample's gate is on
                rows = {"ftab": d.deform_on, "tex": d.deform_on, "surf": d.deform_on, "btab": d.bias_on, "pos": d.res_on, "ztab": d.res_on, "n_out": d.res_on,
                        "ntaps": d.res_on}.get(f, np.ones(B, dtype=bool))
                assert np.array_equal(got[f][rows], S_np[f][rows]), (f, got[f], S_np[f], step)
            import ctypes as C

            for b in range(B):
                nl = d.mus.shape[1]
                assert np.array_equal(np.ctypeslib.as_array(C.cast(int(got["mus"][b]), C.POINTER(C.c_float)), (nl,)), d.mus[b])
                for a in range(3):
                    if not d.res_on[b]:
                        continue
                    pa, pb, n = int(got["taps"][b, a]), int(S_np["taps"][b, a]), int(got["ntaps"][b, a])
                    assert (pa == 0) == (pb == 0)
                    if pa:
                        assert np.array_equal(np.ctypeslib.as_array(C.cast(pa, C.POINTER(C.c_float)), (n,)), np.ctypeslib.as_array(C.cast(pb, C.POINTER(C.c_float)), (n,)))
                        for p in range(a):  # the same width shares one array in both
                            assert (got["taps"][b, p] == got["taps"][b, a]) == (S_np["taps"][b, p] == S_np["taps"][b, a])
    finally:
        restore()
