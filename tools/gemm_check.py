"""GPU check of the tcgen05 GEMM path of matmulDispatch (batch >= 16) against the CPU oracle."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nanollama_b200 import gguf as G, model as M
from oracle import oracle as O

def run(typ, rows, cols, batch, seed=0):
    rng = np.random.default_rng(seed)
    w = (rng.standard_normal((rows, cols)) / np.sqrt(cols)).astype(np.float32)
    raw = G.encode_tensor(w, typ)
    x = rng.standard_normal((batch, cols)).astype(np.float32)
    dm = M.DeviceMatrix(raw, typ, rows, cols)
    got = dm.matmul(x)
    worst = 0.0
    for b in list(range(min(batch, 4))) + [batch - 1]:
        exp = O.matmul(raw, typ, x[b], rows, cols)
        worst = max(worst, float(np.abs(got[b] - exp).max() / np.abs(exp).max()))
    dm.close()
    return worst

if __name__ == "__main__":
    for typ, name in ((G.GGML_Q4_0, "q4_0"), (G.GGML_Q8_0, "q8_0"), (G.GGML_F16, "f16")):
        # batch <= 128: the tall orientation (weights on the M side, split K when the matrix is narrow); above: the wide one
        for rows, cols, batch in ((128, 64, 128), (256, 256, 128), (1000, 768, 200), (4096, 1536, 512), (300, 128, 16), (1536, 4096, 17), (2304, 1536, 64),
                                 (520, 2048, 100), (777, 1024, 129), (4096, 4096, 300), (520, 1376, 40), (300, 1376, 150)):
            try:
                print(name, rows, cols, batch, "max-rel", "%.3e" % run(typ, rows, cols, batch), flush=True)
            except Exception as e:
                print(name, rows, cols, batch, "ERROR", e, flush=True)
