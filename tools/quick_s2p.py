"""Ad-hoc device-resident timing of the sam2pairs kernels (development aid; bench.py is the contract)."""
import ctypes as C
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import microcket_b200 as mk

mode = sys.argv[1] if len(sys.argv) > 1 else "flash"
n_groups = int(sys.argv[2]) if len(sys.argv) > 2 else 4_000_000
window = int(sys.argv[3]) << 20 if len(sys.argv) > 3 else 256 << 20
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 5
L = mk.lib()
n = C.c_size_t()
m = {"flash": 0, "unc": 1}[mode]
rmdup = os.environ.get("MK_RMDUP", "0") == "1"
opts = mk.synth_opts(dup_per_1024=128, dup_universe=n_groups) if rmdup else None
o = C.byref(opts) if opts is not None else None
L.check(L.L.mk_synth_device_ex(0, 1, m, 0, o, 0, n_groups, None, 0, C.byref(n), None))
nbytes = n.value
buf = torch.empty(nbytes + 64, dtype=torch.uint8, device="cuda")
t0 = time.time()
L.check(L.L.mk_synth_device_ex(0, 1, m, 0, o, 0, n_groups, buf.data_ptr(), nbytes, C.byref(n), None))
print(f"generated {nbytes/1e9:.2f} GB in {time.time()-t0:.2f}s")
text = torch.empty(n_groups * 96, dtype=torch.uint8, device="cuda")
pairs = torch.empty(n_groups * 16, dtype=torch.uint8, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
for it in range(iters):
    s = mk.Sam2Pairs(mk.S2PConfig(mode=mode, threads=8, write_sam=False, emit_packed=True, window_bytes=window, rmdup=rmdup, rmdup_capacity=n_groups))
    s.enable_timing()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    io = s.run_device(buf.data_ptr(), nbytes, True, text.data_ptr(), text.numel(), pairs.data_ptr(), n_groups, stream=stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    st = s.finish()
    print(f"iter {it}: {ms:.2f} ms  {nbytes/ms/1e6:.1f} GB/s  {st.pairs/ms/1e3:.2f} Mpairs/s  groups {st.groups} pairs {st.pairs} launches {s.launches()}")
    print({k: round(v[0], 3) for k, v in s.kernel_times().items()}, (s.rmdup_stats().log_text() if rmdup else b""))
    s.close()
