import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mnasnet-pytorch_b200"))
import contextlib, io, torch
from mnb200 import engine
from models.classifiers import FineTuneModelPool, load_model
def run(slack, steps=12):
    torch.manual_seed(42)
    with contextlib.redirect_stdout(io.StringIO()):
        m = FineTuneModelPool(load_model('mnasnet'), 'mnasnet', 1000, '512')
    engine.configure(m, dtype="bf16")
    m = m.cuda().train()
    eng = engine.engine_for(m)
    eng.wgrad_slack = slack
    x = torch.randn(256, 3, 224, 224, device="cuda"); t = torch.randint(0, 1000, (256,), device="cuda")
    for _ in range(4): eng.train_step(x, t)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): eng.train_step(x, t)
    e1.record(); torch.cuda.synchronize()
    print(f"wgrad_slack={slack}: {e0.elapsed_time(e1)/steps:.2f} ms/step", flush=True)
    del eng, m
    torch.cuda.empty_cache()
for s in (0, 2, 4, 8, 12):
    run(s)
