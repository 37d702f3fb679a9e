#!/usr/bin/env python
"""Where the host path of configs[1] loses time against one device-resident launch: the same 4096
frames evaluated (a) in one call, (b) in the host path's chunks on one stream, (c) on two streams."""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cvpack_b200 as cvpack  # noqa: E402
from scripts import workloads  # noqa: E402
frames = 4096
device = torch.device("cuda:0")
cv, system, x, box, _ = workloads.contacts(frames, 100, device)
ctx = cvpack.BatchContext(system, x, box, device=device)
value = torch.empty(frames, device=device)
grad = torch.empty((frames, 20000, 3), device=device)
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
out = {"one_call_ms": timed(lambda: ctx.evaluateInto(cv, value, grad))}
sizes = [296, 876, 876, 876, 876, 296]
subs = []
first = 0
for s in sizes:
    subs.append((cvpack.BatchContext(system, x[first:first + s], box, device=device), value[first:first + s], grad[first:first + s]))
    first += s
def chunks():
    for c, v, g in subs: c.evaluateInto(cv, v, g)
out["chunks_one_stream_ms"] = timed(chunks)
print(json.dumps(out))
# (c) two specs (own workspaces), chunks alternating between two streams
cv2, system2, _, _, _ = workloads.contacts(8, 100, device)
streams = [torch.cuda.Stream(device), torch.cuda.Stream(device)]
subs2 = []
first = 0
for k, s in enumerate(sizes):
    which = k % 2
    c = cvpack.BatchContext(system if which == 0 else system2, x[first:first + s], box, device=device)
    subs2.append((c, cv if which == 0 else cv2, value[first:first + s], grad[first:first + s], streams[which]))
    first += s
def two_streams():
    start = torch.cuda.current_stream(device)
    for st in streams: st.wait_stream(start)
    for c, spec, v, g, st in subs2:
        with torch.cuda.stream(st):
            c.evaluateInto(spec, v, g)
    for st in streams: start.wait_stream(st)
out["chunks_two_streams_ms"] = timed(two_streams)
print(json.dumps(out))
