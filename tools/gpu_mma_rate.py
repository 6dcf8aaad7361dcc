"""MMA issue-rate probe (see experiments.cu::mma_rate_kernel): cycles per tcgen05.mma by shape / operand placement."""
import ctypes, json, sys
import torch
sys.path.insert(0, "/root/repo")
from mangatranslator_b200 import _lib

lib = _lib.exp_lib()
main = _lib.lib()
fn = lib.mtb_exp_mma_rate
fn.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 6 + [ctypes.c_void_p]
fn.restype = ctypes.c_int
out = {}
names = {0: "M128N64", 1: "M128N128", 2: "M128N256", 3: "M128N128+M128N64", 4: "M128N240", 5: "M64N240", 6: "M64N256", 7: "M64N128", 8: "M128N240+M64N240", 9: "M64N64"}
for ctas in (148,):
    for pattern in range(10):
        for (sbo, shift, off) in ((1024, 1024, 0), (1280, 128, 0)):
            cyc = torch.zeros(ctas, dtype=torch.int64, device="cuda")
            iters = 200
            for _ in range(2):
                rc = fn(cyc.data_ptr(), ctas, pattern, iters, sbo, shift, off, _lib.stream_ptr())
                assert rc == 0
                torch.cuda.synchronize()
            n_mma = iters * 36 * (2 if pattern in (3, 8) else 1)
            c = cyc.float()
            key = f"ctas{ctas}_{names[pattern]}_sbo{sbo}_shift{shift}"
            out[key] = {"clk_per_mma_mean": round(float(c.mean()) / n_mma, 2), "max": round(float(c.max()) / n_mma, 2)}
            print(key, out[key], flush=True)
json.dump(out, open("/root/repo/gpurun_out/mma_rate.json", "w"), indent=1)
