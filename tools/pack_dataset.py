"""One-time converter: BIDS segmentations + seed NIfTIs -> bit-packed subject cache files
(fetalsyngen_b200/data/packed.py), the format ``FetalSynthDataset(packed_cache=...)`` loads.

    python tools/pack_dataset.py --bids_path /path/to/bids --seed_path /path/to/seeds --out_path /path/to/cache
"""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

from fetalsyngen_b200.data.datasets import FetalSynthDataset  # noqa: E402
from fetalsyngen_b200.data.packed import pack_subject  # noqa: E402


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("--bids_path", required=True)
    ap.add_argument("--seed_path", required=True)
    ap.add_argument("--out_path", required=True)
    a = ap.parse_args(argv)
    ds = FetalSynthDataset(a.bids_path, None, a.seed_path, None)  # discovery only: no generator, no GPU
    total = 0
    for idx, (sub, ses) in enumerate(ds.sub_ses):
        name = ds._sub_ses_string(sub, ses)
        f = pack_subject(ds.segm_paths[idx], ds.seed_paths[name], Path(a.out_path) / f"{name}.fsgpack.npz")
        total += f.stat().st_size
        print(f"{name}: {f.stat().st_size / 2**20:.1f} MiB")
    print(f"{len(ds.sub_ses)} subjects, {total / 2**20:.1f} MiB")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
