"""Reference-shaped API on the GPU: error behaviour, gates, default probabilities, dtype and
ownership contracts (model.py:94-276, datasets.py:256-349), host pipeline, larger volumes."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
from fetalsyngen_b200 import _lib  # noqa: E402
from fetalsyngen_b200.utils.phantom import label_phantom  # noqa: E402
from gpu_util import DEV  # noqa: E402

pytestmark = pytest.mark.gpu


def _gen(shape, probs=1.0, artifacts=None):
    from fetalsyngen_b200.generator.augmentation.synthseg import RandBiasField, RandGamma, RandNoise, RandResample
    from fetalsyngen_b200.generator.deformation.affine_nonrigid import SpatialDeformation
    from fetalsyngen_b200.generator.intensity.rand_gmm import ImageFromSeeds
    from fetalsyngen_b200.generator.model import FetalSynthGen

    labels = [0] + list(range(10, 50))
    classes = [0] + [10] * 10 + [20] * 10 + [30] * 10 + list(range(40, 50))
    return FetalSynthGen(shape=list(shape), resolution=[0.5] * 3, device=DEV, intensity_generator=ImageFromSeeds(1, 6, labels, classes),
                         spatial_deform=SpatialDeformation(20, 0.02, 0.1, list(shape), probs, True, 0.03, 0.06, 4, 0.5, DEV), resampler=RandResample(probs, 0.5, 1.5),
                         bias_field=RandBiasField(probs, 0.004, 0.02, 0.01, 0.3), noise=RandNoise(probs, 5, 15), gamma=RandGamma(probs, 0.1), **(artifacts or {}))


def test_errors_match_the_reference_contract():
    shape = (32, 32, 32)
    gen = _gen(shape)
    seg = torch.zeros(shape, dtype=torch.float32, device=DEV)
    with pytest.raises(ValueError):  # model.py:132-135: no seeds and no image
        gen.sample(image=None, segmentation=seg, seeds=None)
    with pytest.raises(_lib.FsgError):
        from fetalsyngen_b200.engine import SynthEngine

        SynthEngine(shape, (0.5,) * 3, "cpu")
    seg_h, seeds_h = label_phantom(shape)
    seg_d = torch.from_numpy(seg_h).to(DEV)
    seeds_d = [torch.from_numpy(s).to(DEV) for s in seeds_h]
    with pytest.raises(ValueError):  # more volumes than FSG_MAX_JOBS in one launch
        gen.sample_batch([seg_d] * 17, [seeds_d] * 17)
    with pytest.raises((TypeError, ValueError)):  # seed volumes must be int8/uint8 label maps
        gen.sample_batch([seg_d], [[s.float() for s in seeds_d]])
    with pytest.raises((TypeError, ValueError, _lib.FsgError)):  # segmentation of another shape
        gen.sample_batch([seg_d[:16]], [seeds_d])


def test_all_gates_off_is_identity_on_the_segmentation_and_pure_gmm_on_the_image():
    shape = (32, 32, 32)
    gen = _gen(shape, probs=0.0)
    seg_h, seeds_h = label_phantom(shape)
    seg_d = torch.from_numpy(seg_h).to(DEV)
    seeds_d = [torch.from_numpy(s).to(DEV) for s in seeds_h]
    np.random.seed(3)
    torch.manual_seed(3)
    img, seg, params = gen.sample_batch([seg_d], [seeds_d], scale=False)
    assert torch.equal(seg[0], seg_d)
    p = params[0]
    assert p["deform_params"]["affine"] is None and p["gamma_params"]["gamma"] is None and p["resample_params"]["spacing"] is None and p["noise_params"]["noise_std"] is None
    lab = sum(s.astype(np.int64) for s in seeds_h)
    x = img[0].cpu().numpy()
    assert x.min() >= 0 and np.isfinite(x).all()
    # per-label moments of the Philox GMM draw (the reference's sigma per label is in the drawn table)
    mus, sigmas = gen.intensity_generator.draw_gmm({})  # shapes only: the actual tables are not returned on this path
    assert mus.shape == (50,) and sigmas.shape == (50,)
    for lbl in np.unique(lab)[:6]:
        v = x[lab == lbl]
        if v.size > 500:
            assert 0 < v.std() < 40 and 0 <= v.mean() < 330


def test_default_probabilities_give_mixed_batches():
    """prob 0.9 / flip 0.5 as in the reference's default YAML: some samples skip stages; every volume
    stays finite, in [0,1] after ScaleIntensity, labels stay in the input label set."""
    shape = (64, 64, 64)
    gen = _gen(shape, probs=0.6)
    seg_h, seeds_h = label_phantom(shape)
    seg_d = torch.from_numpy(seg_h).to(DEV)
    seeds_d = [torch.from_numpy(s).to(DEV) for s in seeds_h]
    seen = {"deform_off": 0, "resample_off": 0, "bias_off": 0}
    for it in range(3):
        ids = list(range(8 * it, 8 * it + 8))
        img, seg, params = gen.sample_batch([seg_d] * 8, [seeds_d] * 8, scale=True, sample_ids=ids, base_seed=5)
        assert torch.isfinite(img).all() and float(img.min()) >= 0 and float(img.max()) <= 1
        assert set(torch.unique(seg).tolist()) <= set(np.unique(seg_h).tolist())
        for p in params:
            seen["deform_off"] += p["deform_params"]["affine"] is None
            seen["resample_off"] += p["resample_params"]["spacing"] is None
            seen["bias_off"] += p["bf_params"]["bf_size"] is None
    assert all(v > 0 for v in seen.values()), seen


def test_sample_keeps_dtypes_and_parameter_keys():
    shape = (32, 32, 32)
    gen = _gen(shape)
    seg_h, seeds_h = label_phantom(shape)
    for dt in (torch.float32, torch.int64, torch.uint8):
        seg = torch.from_numpy(seg_h).to(DEV).to(dt)
        plan_seeds = [torch.from_numpy(s).to(DEV) for s in seeds_h]
        img, sg, params = gen.sample_batch([seg.to(torch.uint8)], [plan_seeds])
        assert sg.dtype == torch.uint8 and img.dtype == torch.float32
    for key in ("selected_seeds", "seed_intensities", "deform_params", "gamma_params", "bf_params", "resample_params", "noise_params"):
        assert key in params[0]
    assert set(params[0]["deform_params"]) == {"affine", "non_rigid", "flip"}
    assert set(params[0]["deform_params"]["affine"]) == {"rotations", "shears", "scalings"}


def test_host_pipeline_matches_device_path_and_keeps_order():
    from fetalsyngen_b200.host_pipeline import HostPipeline

    shape = (64, 64, 64)
    gen = _gen(shape)
    seg_h, seeds_h = label_phantom(shape)
    B = 2
    hp = HostPipeline(gen, B, depth=2)
    hp.set_inputs([seg_h] * B, [seeds_h] * B)
    got = []
    steps = 5
    for it in range(steps):
        if it >= 2:
            img, seg, _ = hp.collect()
            got.append((img.clone(), seg.clone()))
        hp.submit(sample_ids=[2 * it, 2 * it + 1], base_seed=9)
    while len(got) < steps:
        img, seg, _ = hp.collect()
        got.append((img.clone(), seg.clone()))
    with pytest.raises(IndexError):
        hp.collect()
    seg_d = torch.from_numpy(seg_h).to(DEV)
    seeds_d = [torch.from_numpy(s).to(DEV) for s in seeds_h]
    for it in range(steps):
        img, seg, _ = gen.sample_batch([seg_d] * B, [seeds_d] * B, scale=True, sample_ids=[2 * it, 2 * it + 1], base_seed=9)
        assert torch.equal(img.cpu(), got[it][0]) and torch.equal(seg.cpu(), got[it][1])
    assert hp.h2d_bytes == B * 5 * 64**3 and hp.d2h_bytes == B * 5 * 64**3


def test_host_pipeline_with_packed_seed_inputs_matches_device_path():
    """HostPipeline(packed_counts=...): the host inputs are the bit-packed seed words of the subject cache
    (3 bytes per voxel in instead of 5); results equal `sample_batch` on PackedSeeds for the same sample ids."""
    from fetalsyngen_b200.data.packed import PackedSeeds, pack_seed_volumes
    from fetalsyngen_b200.host_pipeline import HostPipeline

    shape = (64, 64, 64)
    gen = _gen(shape)
    gen.intensity_generator.min_subclusters, gen.intensity_generator.max_subclusters = 1, 4
    per = {}
    seg_h = None
    for n in range(1, 5):
        seg_h, sv = label_phantom(shape, n_sub=(n, n, n, n), seed=n)
        per[n] = {m + 1: sv[m] for m in range(4)}
    words, counts = pack_seed_volumes(per)
    B = 2
    hp = HostPipeline(gen, B, depth=2, packed_counts=counts)
    hp.set_inputs_packed([seg_h] * B, [words] * B)
    assert hp.h2d_bytes == B * 3 * 64**3 and hp.d2h_bytes == B * 5 * 64**3
    got = []
    for it in range(3):
        hp.submit(sample_ids=[2 * it, 2 * it + 1], base_seed=9)
        img, seg, params = hp.collect()
        got.append((img.clone(), seg.clone(), params))
    seg_d = torch.from_numpy(seg_h).to(DEV)
    ps = PackedSeeds(words, counts, DEV)
    drawn = set()
    for it in range(3):
        img, seg, params = gen.sample_batch([seg_d] * B, [ps] * B, scale=True, sample_ids=[2 * it, 2 * it + 1], base_seed=9)
        assert torch.equal(img.cpu(), got[it][0]) and torch.equal(seg.cpu(), got[it][1])
        for b in range(B):
            m2s = params[b]["selected_seeds"]["mlabel2subclusters"]
            assert m2s == got[it][2][b]["selected_seeds"]["mlabel2subclusters"]
            drawn |= set(m2s.values())
    assert drawn <= {1, 2, 3, 4} and len(drawn) >= 2
    with pytest.raises(RuntimeError):
        HostPipeline(gen, B, depth=1).set_inputs_packed([seg_h] * B, [words] * B)


def test_full_sample_with_all_artifacts_runs_at_96():
    shape = (96, 96, 96)
    gen = _gen(shape, artifacts=bench.default_artifacts(1.0))
    seg_h, seeds_h = label_phantom(shape)
    seg_d = torch.from_numpy(seg_h).to(DEV).float()
    gen.intensity_generator._cache = {}
    np.random.seed(1)
    torch.manual_seed(1)
    # sample() takes the reference's seed-path dictionary; feed decoded volumes through the cache
    seeds = {n: {m: f"mem://{n}/{m}" for m in range(1, 5)} for n in range(1, 7)}
    for n in range(1, 7):
        for m in range(1, 5):
            gen.intensity_generator._cache[(f"mem://{n}/{m}", str(torch.device(DEV)))] = torch.from_numpy(seeds_h[m - 1]).to(DEV)
    out, seg, image, params = gen.sample(image=None, segmentation=seg_d, seeds=seeds)
    assert out.shape == shape and seg.dtype == torch.float32 and image is None and torch.isfinite(out).all()
    assert set(params["artifacts"]) == {"blur_cortex", "struct_noise", "simulate_motion", "boundaries"}
    assert params["artifacts"]["simulate_motion"]["nstacks"] >= 2
    assert float((out == 0).float().mean()) > 0.05  # boundaries masked the background


def test_sample_with_artifacts_is_stream_agnostic_and_reproducible():
    """The artifact path overlaps host and device work (non-blocking uploads, the mask stacks of
    SimulateMotion on a side stream): the same seeds must give the same sample on the default stream and
    on a private stream, call after call."""
    shape = (64, 64, 64)
    gen = _gen(shape, artifacts=bench.default_artifacts(1.0))
    seg_h, seeds_h = label_phantom(shape)
    seg_d = torch.from_numpy(seg_h).to(DEV).float()
    seeds = {n: {m: f"mem://{n}/{m}" for m in range(1, 5)} for n in range(1, 7)}
    for n in range(1, 7):
        for m in range(1, 5):
            gen.intensity_generator._cache[(f"mem://{n}/{m}", str(torch.device(DEV)))] = torch.from_numpy(seeds_h[m - 1]).to(DEV)

    def run(stream):
        np.random.seed(4)
        torch.manual_seed(4)
        gen._sample_counter = 0
        with torch.cuda.stream(stream):
            out, seg, _, params = gen.sample(image=None, segmentation=seg_d, seeds=seeds)
            stream.synchronize()
        return out.clone(), seg.clone(), params["artifacts"]["simulate_motion"]["nstacks"]

    private = torch.cuda.Stream(device=DEV)
    ref = run(torch.cuda.current_stream())
    for stream in (private, torch.cuda.current_stream(), private):
        got = run(stream)
        assert got[2] == ref[2]
        assert torch.equal(got[1], ref[1])
        # the PSF reconstruction accumulates with float atomics: equal up to summation order
        assert float((got[0] - ref[0]).abs().max()) <= 1e-4 * float(ref[0].max() - ref[0].min())


def test_volume_384_single_sample():
    shape = (384, 384, 384)
    gen = _gen(shape)
    seg_h, seeds_h = label_phantom(shape)
    seg_d = torch.from_numpy(seg_h).to(DEV)
    seeds_d = [torch.from_numpy(s).to(DEV) for s in seeds_h]
    img, seg, _ = gen.sample_batch([seg_d], [seeds_d], scale=True, sample_ids=[0], base_seed=1)
    assert img.shape == (1, *shape) and torch.isfinite(img).all() and float(img.max()) == 1.0
    assert set(torch.unique(seg).tolist()) <= set(np.unique(seg_h).tolist())


def test_device_batch_loader_matches_sample_batch_and_keeps_batches_valid(tmp_path):
    """DeviceBatchLoader (the GPU-resident DataLoader replacement): batches equal the ones
    `generator.sample_batch` produces for the same sample ids, stay on the device, and a yielded batch
    is still intact after the next one has been requested (depth 3: valid for depth - 1 further requests)."""
    from fetalsyngen_b200.data.datasets import DeviceBatchLoader, FetalSynthDataset
    from fetalsyngen_b200.sharding import step_ids
    from fetalsyngen_b200.utils import nifti

    shape = (64, 64, 64)
    seg_h, seeds_h = label_phantom(shape)
    aff = np.diag([0.5, 0.5, 0.5, 1.0])
    for sub in ("sub-a", "sub-b"):
        d = tmp_path / "bids" / sub / "anat"
        d.mkdir(parents=True)
        nifti.write_nifti(d / f"{sub}_rec-x_T2w_dseg.nii.gz", seg_h.astype(np.float32), aff)
        for n in range(1, 3):
            sd = tmp_path / "seeds" / f"subclasses_{n}" / sub / "anat"
            sd.mkdir(parents=True)
            for m in range(1, 5):
                nifti.write_nifti(sd / f"{sub}_rec-x_T2w_dseg_mlabel_{m}.nii.gz", seeds_h[m - 1], aff)
    gen = _gen(shape)
    gen.intensity_generator.min_subclusters, gen.intensity_generator.max_subclusters = 1, 2
    ds = FetalSynthDataset(str(tmp_path / "bids"), gen, str(tmp_path / "seeds"), None)
    loader = DeviceBatchLoader(ds, batch_size=2, num_batches=4, shuffle=False, labels_int64=True, base_seed=21, depth=3)
    assert len(loader) == 4
    kept = []
    for step, batch in enumerate(loader):
        assert batch["image"].shape == (2, 1, *shape) and batch["image"].device.type == "cuda" and batch["label"].dtype == torch.int64
        assert len(batch["name"]) == 2 and len(batch["params"]) == 2
        kept.append((batch["image"], batch["label"], batch["image"].clone(), batch["label"].clone()))
        if step >= 1:  # the previous batch (other slot) must not have been overwritten yet
            torch.cuda.synchronize()
            assert torch.equal(kept[step - 1][0], kept[step - 1][2]) and torch.equal(kept[step - 1][1], kept[step - 1][3])
    torch.cuda.synchronize()
    # same ids through the plain batched call
    for step in range(4):
        idx = [(step * 2 + k) % len(ds) for k in range(2)]
        segs = [ds._segmentation(i) for i in idx]
        names = [ds._sub_ses_string(*ds.sub_ses[i]) for i in idx]
        img, seg, _ = gen.sample_batch(segs, [ds.seed_paths[n] for n in names], scale=True, sample_ids=step_ids(step, 2, 0, 1), base_seed=21)
        assert torch.equal(img.unsqueeze(1), kept[step][2]) and torch.equal(seg.unsqueeze(1).long(), kept[step][3])
    with pytest.raises(ValueError):
        DeviceBatchLoader(ds, batch_size=2, depth=1)


# ----------------------------------------------------------------------------- bit-packed seed cache
@pytest.mark.parametrize("smax,nvox_shape", [(6, (24, 20, 28)), (10, (24, 20, 28)), (6, (5, 7, 3)), (3, (1, 1, 13))])
def test_unpack_seeds_kernel_equals_sum_of_selected_seed_files(smax, nvox_shape):
    """fsg_unpack_seeds (uint16 and uint32 words, vector body + scalar tail) against the reference's
    definition of the label map: the sum of the four selected seed volumes (rand_gmm.py:90-97)."""
    from fetalsyngen_b200.data import packed as K

    seeds = {}
    for n in range(1, smax + 1):
        _, sv = label_phantom(nvox_shape, n_sub=(n, n, n, n), seed=n)
        seeds[n] = {m + 1: sv[m] for m in range(4)}
    words, counts = K.pack_seed_volumes(seeds)
    ps = K.PackedSeeds(words, counts, DEV)
    rs = np.random.RandomState(1)
    for _ in range(6):
        m2s = {m: int(rs.randint(1, smax + 1)) for m in range(1, 5)}
        want = sum(seeds[m2s[m]][m].astype(np.int32) for m in range(1, 5)).astype(np.uint8)
        got = ps.labels(m2s, DEV).cpu().numpy()
        assert np.array_equal(got, want), m2s
        assert np.array_equal(got, K.unpack_numpy(words, counts, m2s))
    with pytest.raises(KeyError):
        ps.labels({1: smax + 1, 2: 1, 3: 1, 4: 1}, DEV)


def test_dataset_with_packed_cache_gives_identical_samples(tmp_path):
    """FetalSynthDataset(packed_cache=...) converts on first use, reloads from the cache file afterwards,
    and produces bit-identical samples to the NIfTI-backed dataset (per-sample API and batched path)."""
    from fetalsyngen_b200.data.datasets import FetalSynthDataset
    from fetalsyngen_b200.utils import nifti

    shape = (48, 48, 48)
    aff = np.diag([0.5, 0.5, 0.5, 1.0])
    for si, sub in enumerate(("sub-a", "sub-b")):
        d = tmp_path / "bids" / sub / "anat"
        d.mkdir(parents=True)
        seg_h, _ = label_phantom(shape, seed=si)
        nifti.write_nifti(d / f"{sub}_rec-x_T2w_dseg.nii.gz", seg_h.astype(np.float32), aff)
        for n in range(1, 5):
            _, sv = label_phantom(shape, n_sub=(n, n, n, n), seed=10 * si + n)
            sd = tmp_path / "seeds" / f"subclasses_{n}" / sub / "anat"
            sd.mkdir(parents=True)
            for m in range(1, 5):
                nifti.write_nifti(sd / f"{sub}_rec-x_T2w_dseg_mlabel_{m}.nii.gz", sv[m - 1], aff)
    gen = _gen(shape)
    gen.intensity_generator.min_subclusters, gen.intensity_generator.max_subclusters = 1, 4
    plain = FetalSynthDataset(str(tmp_path / "bids"), gen, str(tmp_path / "seeds"), None)
    cache = tmp_path / "cache"
    for attempt in range(2):  # first pass converts, second pass only reads the cache files
        packed = FetalSynthDataset(str(tmp_path / "bids"), gen, str(tmp_path / "seeds"), None, packed_cache=str(cache))
        for idx in (0, 1):
            outs = []
            for ds in (plain, packed):
                np.random.seed(5 + idx)
                torch.manual_seed(5 + idx)
                gen._sample_counter = 0
                o, p = ds.sample(idx)
                outs.append((o, p))
            (a, pa), (b, pb) = outs
            assert torch.equal(a["image"], b["image"]) and torch.equal(a["label"], b["label"])
            assert pa["selected_seeds"] == pb["selected_seeds"]
        assert sorted(f.name for f in cache.glob("*.npz")) == ["sub-a.fsgpack.npz", "sub-b.fsgpack.npz"]
    a, _ = plain.sample_batch([0, 1, 0])
    np.random.seed(3)
    torch.manual_seed(3)
    ia, sa, _ = gen.sample_batch([plain._segmentation(i) for i in (0, 1)], [plain._seeds(i) for i in (0, 1)], sample_ids=[4, 9], base_seed=77)
    ib, sb, _ = gen.sample_batch([packed._segmentation(i) for i in (0, 1)], [packed._seeds(i) for i in (0, 1)], sample_ids=[4, 9], base_seed=77)
    assert torch.equal(ia, ib) and torch.equal(sa, sb)


def test_packed_only_dataset_and_dataset_pipeline_on_the_bundled_subjects():
    """FetalSynthDataset.from_packed over the committed 256^3 subject files (no BIDS tree, no NIfTI): the index-based
    batched entry point, its host-output pipeline (what bench.py's e2e_dataset_cache leg times) and the
    reference-shaped per-sample call all run from the device-resident subject cache."""
    from fetalsyngen_b200.data.datasets import FetalSynthDataset
    from fetalsyngen_b200.host_pipeline import DatasetPipeline
    from golden_util import SUBJECTS

    shape = (256, 256, 256)
    gen = _gen(shape)
    ds = FetalSynthDataset.from_packed(SUBJECTS, gen)
    assert [ds._sub_ses_idx(i) for i in range(3)] == ["sub-sta21", "sub-sta30", "sub-sta38"]
    out, params = ds.sample_batch([0, 1, 2], scale=True, sample_ids=[10, 11, 12], base_seed=3)
    img, lab = out["image"][:, 0].clone(), out["label"][:, 0].clone()
    assert img.shape == (3, *shape) and float(img.max()) == 1.0 and float(img.min()) == 0.0 and out["name"] == ["sub-sta21", "sub-sta30", "sub-sta38"]
    # same ids through the generator directly: identical volumes (every draw is a function of (base_seed, id))
    segs = [ds._segmentation(i) for i in range(3)]
    img2, lab2, _ = gen.sample_batch(segs, [ds._seeds(i) for i in range(3)], scale=True, sample_ids=[10, 11, 12], base_seed=3)
    assert torch.equal(img, img2) and torch.equal(lab, lab2)
    # warped labels stay inside the subject's label set and cover a plausible part of the volume
    for b in range(3):
        assert set(torch.unique(lab[b]).tolist()) <= set(torch.unique(segs[b]).tolist())
        assert 0.3 < float((lab[b] > 0).float().mean()) / float((segs[b] > 0).float().mean()) < 3.0
    # host-output pipeline: two pipelined steps, results equal to the device path and in submission order
    dp = DatasetPipeline(ds, 3, depth=2)
    got = []
    dp.run([[0, 1, 2], [2, 0, 1]], on_result=lambda hi, hs, pr: got.append((hi.clone(), hs.clone())), sample_ids=None)
    assert len(got) == 2 and got[0][0].shape == (3, *shape) and got[0][0].dtype == torch.float32 and got[0][1].dtype == torch.uint8
    dp2 = DatasetPipeline(ds, 3, depth=2)
    dp2.submit([0, 1, 2], True, sample_ids=[10, 11, 12], base_seed=3)
    hi, hs, _ = dp2.collect()
    assert torch.equal(hi, img.cpu()) and torch.equal(hs, lab.cpu())
    # the reference's per-sample entry point on the same cache
    np.random.seed(0)
    torch.manual_seed(0)
    data, gp = ds.sample(1)
    assert data["image"].shape == (1, *shape) and data["label"].dtype == torch.int64 and data["name"] == "sub-sta30" and "generation_time" in gp


def test_batched_path_with_sr_artifacts_equals_the_per_sample_artifact_calls():
    """sample_batch(..., artifacts=True): the four SR artifacts run on every sample of a generated batch (device-resident,
    their draws reseeded per sample id), ScaleIntensity after them — equal to applying FetalSynthGen._run_artifacts by hand
    to the un-scaled batch with the same seeds, and independent of how the ids are batched."""
    import sys

    sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parents[1]))
    import bench
    from fetalsyngen_b200.sharding import sample_seed

    shape = (96, 96, 96)
    gen = _gen(shape, artifacts=bench.default_artifacts(1.0))
    seg_h, seeds_h = label_phantom(shape)
    seg_d = torch.from_numpy(seg_h).to(DEV)
    seeds_d = [torch.from_numpy(s).to(DEV) for s in seeds_h]
    ids, base = [5, 6, 7], 11
    img, seg, params = gen.sample_batch([seg_d] * 3, [seeds_d] * 3, scale=True, sample_ids=ids, base_seed=base, artifacts=True)
    img, seg = img.clone(), seg.clone()
    assert float(img.max()) == 1.0 and float(img.min()) == 0.0
    assert all(set(params[b]["artifacts"]) == {"blur_cortex", "struct_noise", "simulate_motion", "boundaries"} for b in range(3))
    # by hand: un-scaled base batch, artifacts per sample under the same per-sample seeds, then ScaleIntensity
    base_img, base_seg, _ = gen.sample_batch([seg_d] * 3, [seeds_d] * 3, scale=False, sample_ids=ids, base_seed=base)
    base_img = base_img.clone()
    assert torch.equal(base_seg, seg)
    eng = gen.engine(shape)
    for b in range(3):
        sd = sample_seed(base, ids[b])
        np.random.seed(sd)
        torch.manual_seed(sd)
        out, _ = gen._run_artifacts(base_img[b], base_seg[b], {})
        want = eng.scale_intensity(out.contiguous()).view(shape)
        # the PSF reconstruction accumulates with floating-point atomics: two runs agree to rounding, not bit for bit
        diff = (want - img[b]).abs()
        assert float(diff.mean()) <= 1e-4 and float(torch.quantile(diff.flatten()[::7], 0.999)) <= 2e-3, (b, float(diff.max()), float(diff.mean()))
        assert float((base_img[b] / base_img[b].max() - want).abs().mean()) > 1e-3  # the artifacts did change the volume
    # a different batching of the same ids gives the same volumes
    one, _, _ = gen.sample_batch([seg_d], [seeds_d], scale=True, sample_ids=[6], base_seed=base, artifacts=True)
    diff = (one[0] - img[1]).abs()
    assert float(diff.mean()) <= 1e-4 and float(torch.quantile(diff.flatten()[::7], 0.999)) <= 2e-3, (float(diff.max()), float(diff.mean()))


@pytest.mark.parametrize("shape,probs,B", [((64, 48, 80), 0.7, 16), ((96, 96, 96), 1.0, 5)])
def test_native_step_equals_the_python_builders_on_the_device(shape, probs, B, monkeypatch):
    """The batched path through the native step (fsg_draw_batch + fsg_step_fill + fsg_step_run), through the numpy job
    builder and through the per-sample builder: the same launches, so the same volumes bit for bit — with per-sample
    gates (probabilities < 1), a non-cubic shape and a full batch of FSG_MAX_JOBS samples (label volumes as seeds; the
    packed-seed route is compared the same way in test_packed_only_dataset_and_dataset_pipeline_on_the_bundled_subjects)."""
    import fetalsyngen_b200.generator.model as M
    gen = _gen(shape, probs=probs)
    seg_h, seeds_h = label_phantom(shape)
    seg_d = torch.from_numpy(seg_h).to(DEV)
    vols = [torch.from_numpy(s).to(DEV) for s in seeds_h]
    ids = list(range(40, 40 + B))
    outs = []
    for native, fast in ((True, True), (False, True), (False, False)):
        monkeypatch.setattr(M, "_NATIVE_STEP", native)
        monkeypatch.setattr(M, "_FAST_STEP", fast)
        img, seg, params = gen.sample_batch([seg_d] * B, [vols] * B, scale=True, sample_ids=ids, base_seed=8)
        outs.append((img.clone(), seg.clone(), [params[b]["resample_params"]["spacing"] for b in range(B)]))
    for k in (1, 2):
        assert torch.equal(outs[0][0], outs[k][0]) and torch.equal(outs[0][1], outs[k][1]) and outs[0][2] == outs[k][2], k
    if probs < 1:
        assert any(s is None for s in outs[0][2]) and any(s is not None for s in outs[0][2])  # both kinds of sample in the batch
