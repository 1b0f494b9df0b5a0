"""GPU seed generation with the reference script's flags (scripts/generate_seeds.py):

    python tools/generate_seeds.py --bids_path /path/to/bids --out_path /path/to/out --max_subclasses 6 --annotation feta
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

from fetalsyngen_b200.seeds import main  # noqa: E402

if __name__ == "__main__":
    raise SystemExit(main())
