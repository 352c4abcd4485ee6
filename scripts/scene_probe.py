import os, sys, time, json
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import probav_b200 as pb
from probav_b200 import synth
cfg = pb.parseConfig("cfg/p16t9c85r12.cfg")
m = pb.build_from_config(cfg, precision="tf32")
for ns in (32, 128):
    lr, _, _ = synth.make_scene(ns, seed=3)
    m.predict_from_scenes(lr); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3): m.predict_from_scenes(lr)
    dt = (time.perf_counter() - t0) / 3
    d = torch.from_numpy(lr).cuda()
    m.predict_from_scenes(d); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): m.predict_from_scenes(d)
    e1.record(); torch.cuda.synchronize()
    print(json.dumps({"ns": ns, "host_scenes_per_s": ns / dt, "device_scenes_per_s": ns * 3 / (e0.elapsed_time(e1) * 1e-3)}))
