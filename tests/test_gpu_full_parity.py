"""Voxel-wise parity at BASELINE.json's full size (256^3) on the bundled subjects (configs[0] / configs[1]):

* the CUDA base path against the goldens of the UNMODIFIED reference (tests/golden/full_*.npz: warped
  segmentation over the whole volume, the image on a strided sample, per-plane sums);
* the CUDA base path against the numpy oracle, every voxel, for further random draws;
* the batched production call (`run_base` with B > 1, Philox noise replaced by injected tensors) equal to the
  same samples generated one at a time.

Inputs are the committed subject fixtures (tests/golden/subjects: the reference's sub-sta21/30/38, bit-packed);
the label volume of a draw is produced on the device by `fsg_unpack_seeds` (checked against `unpack_numpy`)."""
import numpy as np
import pytest
import torch

import np_oracle as O
from fetalsyngen_b200.data.packed import PackedSeeds, unpack_numpy
from fetalsyngen_b200.engine import engine_for
from golden_util import FULL_CASES, SUB, load_full_case, load_subject, seeded_normal
from gpu_util import DEV, TOL, plan_from_golden

pytestmark = pytest.mark.gpu
RES = (0.5, 0.5, 0.5)


def _device_labels(subject, m2s):
    seg, words, counts = load_subject(subject)
    ps = PackedSeeds(words, counts, device=DEV)
    lab = ps.labels(m2s, DEV)
    return torch.from_numpy(seg).to(DEV).view(-1), lab.view(-1), (seg, words, counts)


def _compare(img, seg, want_img, want_seg):
    assert np.array_equal(seg.cpu().numpy().reshape(want_seg.shape), want_seg), "warped segmentation is not bit-exact"
    rng = float(want_img.max() - want_img.min())
    err = float(np.abs(img.cpu().numpy().reshape(want_img.shape).astype(np.float64) - want_img).max()) / rng
    assert err <= TOL, err
    return err


@pytest.mark.parametrize("name", FULL_CASES)
def test_base_path_vs_reference_and_oracle_at_256(name):
    d, labels, seg_in, p = load_full_case(name)
    m2s = {m: int(d["mlabel2subclusters"][m - 1]) for m in range(1, 5)}
    seg_d, lab_d, _ = _device_labels(str(d["subject"]), m2s)
    assert np.array_equal(lab_d.cpu().numpy().reshape(labels.shape), labels)  # fsg_unpack_seeds, bit-exact
    eng = engine_for(DEV, tuple(seg_in.shape), RES)
    plan = plan_from_golden(d)
    img, sg = eng.run_base([plan], [[lab_d]], [seg_d])
    torch.cuda.synchronize()
    got, got_seg = img[0].cpu().numpy(), sg[0].cpu().numpy()
    # ---- the unmodified reference
    assert np.array_equal(got_seg, d["seg_out"]), "segmentation differs from the reference at 256^3"
    rng = float(d["final_max"] - d["final_min"])
    assert np.abs(got[SUB].astype(np.float64) - d["final_sub"]).max() / rng <= TOL
    nplane = got.shape[1] * got.shape[2]
    assert np.abs(got.astype(np.float64).sum(axis=(1, 2)) - d["final_plane_sums"]).max() / nplane / rng <= TOL
    # ---- the oracle, every voxel
    want, want_seg, _ = O.generate_base(labels, seg_in, p)
    _compare(img[0], sg[0], want, want_seg)
    # ---- ScaleIntensity fused into the last kernel (datasets.py:311)
    img2, _ = eng.run_base([plan], [[lab_d]], [seg_d], scale=True)
    ws = O.scale_intensity(want)
    assert float(np.abs(img2[0].cpu().numpy().astype(np.float64) - ws).max()) <= TOL
    assert float(img2.max()) == 1.0 and float(img2.min()) == 0.0


def _random_params(rs, shape, flip, resample):
    """One draw in np_oracle.generate_base form with seeded volume noise (default config ranges)."""
    from fetalsyngen_b200.tables import make_affine_matrix, resample_size, resample_stds

    q = {"mus": (25 + 200 * rs.rand(50)).astype(np.float32), "sigmas": (5 + 20 * rs.rand(50)).astype(np.float32), "flip": flip, "resolution": np.array(RES), "size": shape}
    q["gmm_noise"] = seeded_normal(rs.randint(1 << 30), shape)
    rot = (2 * 20 * rs.rand(3) - 20) / 180 * np.pi
    q["A"] = make_affine_matrix(rot, 0.04 * rs.rand(3) - 0.02, 1 + 0.2 * rs.rand(3) - 0.1).astype(np.float32)
    q["c2"] = (np.array(shape) - 1) / 2
    s = [int(round((0.03 + 0.03 * rs.rand()) * v)) for v in shape]
    q["Fsmall"] = (4 * rs.rand() * rs.randn(*s, 3)).astype(np.float32)
    q["gamma"] = float(np.exp(0.1 * rs.randn()))
    b = [max(int(round((0.004 + 0.016 * rs.rand()) * v)), 1) for v in shape]
    q["bf_low"] = ((0.01 + 0.29 * rs.rand()) * rs.randn(*b)).astype(np.float32)
    q["noise_std"] = float(5 + 10 * rs.rand())
    if resample:
        sp = 0.5 + rs.rand()
        q["spacing"] = np.array([sp] * 3)
        q["stds"] = resample_stds(q["spacing"], RES, rs.rand())
        q["noise"] = seeded_normal(rs.randint(1 << 30), [resample_size(v, 0.5, sp) for v in shape])
    else:
        q["noise"] = seeded_normal(rs.randint(1 << 30), shape)
    return q


def _plan_from_params(q):
    from fetalsyngen_b200.engine import SamplePlan

    p = SamplePlan(mus=q["mus"], sigmas=q["sigmas"], gmm_noise=torch.from_numpy(q["gmm_noise"]).to(DEV).view(-1))
    p.deform, p.flip, p.A, p.c2, p.fsmall = True, q["flip"], q["A"], q["c2"], q["Fsmall"]
    p.center = ((np.array(q["size"]) - 1) / 2).astype(np.float32)
    p.gamma, p.bf_low, p.noise_std = q["gamma"], q["bf_low"], q["noise_std"]
    p.noise = torch.from_numpy(q["noise"]).to(DEV).view(-1)
    if "spacing" in q:
        p.spacing, p.stds = q["spacing"], q["stds"]
    return p


def test_batched_base_path_vs_oracle_every_voxel_at_256():
    """Three subjects, three draws (flip / no flip / no resolution simulation) in ONE batched call: every voxel
    of every sample against the oracle; this is the launch shape bench.py times."""
    rs = np.random.RandomState(2024)
    cases = [("sub-sta21", True, True), ("sub-sta30", False, True), ("sub-sta38", True, False)]
    plans, labs, segs, want = [], [], [], []
    for subject, flip, resample in cases:
        m2s = {m: int(rs.randint(1, 7)) for m in range(1, 5)}
        seg_d, lab_d, (seg, words, counts) = _device_labels(subject, m2s)
        q = _random_params(rs, tuple(seg.shape), flip, resample)
        plans.append(_plan_from_params(q))
        labs.append([lab_d])
        segs.append(seg_d)
        want.append(O.generate_base(unpack_numpy(words, counts, m2s), seg, q)[:2])
    eng = engine_for(DEV, (256, 256, 256), RES)
    img, sg = eng.run_base(plans, labs, segs)
    torch.cuda.synchronize()
    for b in range(len(cases)):
        _compare(img[b], sg[b], *want[b])


def test_gmm_reads_packed_seed_words_directly():
    """fsg_gmm's packed mode (labels decoded in the kernel from the subject's bit-packed words: what the dataset
    cache path runs) against the unpack-then-sum route: label volume and Philox intensities bit-identical."""
    from fetalsyngen_b200.engine import SamplePlan

    rs = np.random.RandomState(9)
    seg, words, counts = load_subject("sub-sta21")
    ps = PackedSeeds(words, counts, device=DEV)
    eng = engine_for(DEV, tuple(seg.shape), RES)
    plans, sel = [], []
    for b in range(3):
        plans.append(SamplePlan(mus=(25 + 200 * rs.rand(50)).astype(np.float32), sigmas=(5 + 20 * rs.rand(50)).astype(np.float32), rng_seed=5, sample_id=b))
        sel.append({m: int(rs.randint(1, 7)) for m in range(1, 5)})
    out_a = torch.empty((3, eng.nvox), dtype=torch.float32, device=DEV)
    lab_a = torch.empty((3, eng.nvox), dtype=torch.uint8, device=DEV)
    eng.gmm(plans, [(ps, m2s) for m2s in sel], out_a, labels_out=lab_a)
    out_b, lab_b = torch.empty_like(out_a), torch.empty_like(lab_a)
    eng.gmm(plans, [[ps.labels(m2s, DEV).view(-1)] for m2s in sel], out_b, labels_out=lab_b)
    for b in range(3):
        assert np.array_equal(lab_a[b].cpu().numpy().reshape(seg.shape), unpack_numpy(words, counts, sel[b]))
    assert torch.equal(lab_a, lab_b) and torch.equal(out_a, out_b)


def test_texture_gather_hand_over_is_bit_identical_at_256():
    """The default GMM -> warp hand-over (block-linear fsg_texvol written through a surface, gathered with tld4)
    against the linear float32 hand-over (FSG_WARP_TEX=0): image and segmentation bit for bit, for flipped and
    unflipped samples with Philox noise in one batched launch; the GMM kernel's surface output equals its linear
    output."""
    from fetalsyngen_b200.engine import SamplePlan, TexVolume

    rs = np.random.RandomState(77)
    eng = engine_for(DEV, (256, 256, 256), RES)
    if __import__("os").environ.get("FSG_WARP_TEX", "1") == "0":
        pytest.skip("texture hand-over switched off (FSG_WARP_TEX=0)")
    assert eng.use_tex, "the texture hand-over is the default"
    plans, seeds, segs = [], [], []
    for b, (subject, flip, resample) in enumerate([("sub-sta21", True, True), ("sub-sta30", False, True), ("sub-sta38", True, False)]):
        seg, words, counts = load_subject(subject)
        q = _random_params(rs, tuple(seg.shape), flip, resample)
        p = _plan_from_params(q)
        p.gmm_noise, p.noise, p.rng_seed, p.sample_id = None, None, 123, b
        assert eng.tex_eligible(p)
        plans.append(p)
        seeds.append((PackedSeeds(words, counts, device=DEV), {m: int(rs.randint(1, 7)) for m in range(1, 5)}))
        segs.append(torch.from_numpy(seg).to(DEV).view(-1))
    img_t, seg_t = eng.run_base(plans, seeds, segs)
    eng.use_tex = False
    try:
        img_l, seg_l = eng.run_base(plans, seeds, segs)
    finally:
        eng.use_tex = True
    assert torch.equal(seg_t, seg_l) and torch.equal(img_t, img_l)
    # the GMM stage alone: surface output read back == linear output
    lin = torch.empty((3, eng.nvox), dtype=torch.float32, device=DEV)
    eng.gmm(plans, seeds, lin)
    vols = [TexVolume(eng.shape) for _ in range(3)]
    eng.gmm(plans, seeds, [None] * 3, tex=vols)
    back = torch.empty_like(lin)
    for b in range(3):
        vols[b].download(back[b])
    torch.cuda.synchronize()
    assert torch.equal(back, lin)
