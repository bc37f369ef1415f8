import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wobble_jax_b200 import engine, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
f32 = len(sys.argv) > 2 and sys.argv[2] == "f32"
plans, ps, refs = [], [], []
for c in range(n):
    ch = synth.make_chunk(c, as_block=True)
    model = synth.fit_all(ch.model)
    plan = engine.Plan(engine.TreeSpec(model), ch.datablock["xs"], ch.datablock["ys"], ch.datablock["yivar"], ch.datablock["mask"], ch.metablock, float32=f32)
    plan.sync_from_model()
    p = np.asarray(model.get_parameters()); ps.append(p); plans.append(plan)
    refs.append(plan.eval(p))
d_p = [torch.from_numpy(p.copy()).cuda() for p in ps]
d_out = [torch.zeros(len(p) + 1, dtype=torch.float64, device="cuda") for p in ps]
batch = engine.Batch(plans)
batch.bind([t.data_ptr() for t in d_p], [t.data_ptr() for t in d_out])
s_ = torch.cuda.Stream(); s_.wait_stream(torch.cuda.current_stream()); torch.cuda.set_stream(s_)
st = s_.cuda_stream
for _ in range(3): batch.eval_device(st)
torch.cuda.synchronize(); batch.check()
for i, ((f, g), o) in enumerate(zip(refs, d_out)):
    o = o.cpu().numpy()
    print(i, o[0] == f, np.array_equal(o[1:], g), f, o[0])
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): batch.eval_device(st)
e1.record(); torch.cuda.synchronize()
print("ms/step", e0.elapsed_time(e1) / 10, "per chunk us", e0.elapsed_time(e1) / 10 / n * 1e3)
