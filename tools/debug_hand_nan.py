"""Debug aid: poison the allocator's free memory with NaNs, then run the batched hand path and report where
non-finite values first appear."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import isl_b200
from isl_b200 import synth

torch.cuda.set_device(0)
poison = [torch.full((256 * 1024 * 1024,), float("nan"), device="cuda") for _ in range(8)]   # 8 GiB of NaN
del poison
hand = isl_b200.Hand(synth.make_flat_weights("hand", seed=2))
crops = [synth.synth_frame(64, 64, 5), synth.synth_frame(109, 109, 6), synth.synth_frame(64, 64, 7)]
dev = [torch.from_numpy(c).cuda() for c in crops]
per_crop = hand.network_outputs(dev)
torch.cuda.synchronize()
for (key, inst) in hand.model._instances.items():
    out = inst.outputs[0]
    print("instance", key, "output finite per slot:", [bool(torch.isfinite(out[i]).all()) for i in range(out.shape[0])],
          "absmax", [float(out[i].nan_to_num().abs().max()) for i in range(out.shape[0])])
    for name, buf in inst.bufs.items():
        bad = [i for i in range(buf.shape[0]) if not torch.isfinite(buf[i].float()).all()]
        if bad:
            print("   buffer", name, tuple(buf.shape), "non-finite in slots", bad)
# single instance, no concurrency: does the 736 plan produce NaNs by itself on finite input?
inst = hand.model.instance(4, 736, 736)
inst.input.zero_()
inst.input[:3].uniform_(-0.5, 0.5)
inst.run(); torch.cuda.synchronize()
print("alone 736: finite per slot", [bool(torch.isfinite(inst.outputs[0][i]).all()) for i in range(4)],
      {k: [i for i in range(4) if not torch.isfinite(v[i].float()).all()] for k, v in inst.bufs.items() if not torch.isfinite(v.float()).all()})
for size in (736, 552, 184):
    inst = hand.model.instance(4, size, size)
    inst.input[:3].uniform_(-0.5, 0.5)
    inst.input[3] = float("nan")
    inst.run(); torch.cuda.synchronize()
    print("alone %d with NaN slot 3: finite per slot" % size, [bool(torch.isfinite(inst.outputs[0][i]).all()) for i in range(4)],
          {k: [i for i in range(4) if not torch.isfinite(v[i].float()).all()] for k, v in inst.bufs.items()
           if not torch.isfinite(v[:3].float()).all()})
print("peaks", [p[:3].tolist() for p in hand.batch(crops)])
