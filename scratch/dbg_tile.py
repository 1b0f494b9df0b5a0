import sys; sys.path[:0]=['.','oracle','tests']
import numpy as np, torch
from golden_util import load_case
from gpu_util import plan_from_golden, seeds_from_golden, engine_from_golden
d = load_case("c32_all")
eng = engine_from_golden(d); plan = plan_from_golden(d)
seg = torch.from_numpy(d["seg_in"]).cuda().contiguous().view(-1)
img, sg = eng.run_base([plan], [seeds_from_golden(d)], [seg], scale=False)
torch.cuda.synchronize()
print("seg equal", np.array_equal(sg[0].cpu().numpy(), d["seg_out"]))
ref = d["final"] if "final" in d else None
if ref is not None: print("img err", float(np.abs(img[0].cpu().numpy()-ref).max()/(ref.max()-ref.min())))
