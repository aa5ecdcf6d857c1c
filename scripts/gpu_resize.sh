#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_resize_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/k_resize.log 2>&1
echo "== resize tests exit $?" > gpurun_out/summary.txt; tail -15 gpurun_out/k_resize.log >> gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_pipeline_gpu.py -m gpu -q -x -p no:cacheprovider -k "driver" > gpurun_out/k_driver.log 2>&1
echo "== driver test exit $?" >> gpurun_out/summary.txt; tail -5 gpurun_out/k_driver.log >> gpurun_out/summary.txt
timeout 300 python /dev/stdin >> gpurun_out/summary.txt 2>&1 <<'PY'
import sys, time, numpy as np, torch
sys.path.insert(0, ".")
from PIL import Image
from infer import _driver_common as dc
a = np.random.default_rng(0).integers(0, 256, (128, 128, 3), dtype=np.uint8)
img = Image.fromarray(a, "RGB")
for name, fn in (("PIL host chain", lambda: dc.pil_to_model_input(dc.preprocess(img, 512, 4)[0], "cuda", torch.bfloat16)),
                 ("GPU chain", lambda: dc.preprocess_gpu(img, 512, 4, "cuda", torch.bfloat16)[1])):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): fn()
    torch.cuda.synchronize(); print(f"pre-processing 128->512, {name}: {(time.perf_counter() - t0) / 20 * 1e3:.2f} ms/image")
out = (torch.rand(1, 3, 512, 512) * 2 - 1).bfloat16().cuda()
lq_pil = dc.preprocess(img, 512, 4)[0]; lq_u8 = dc.preprocess_gpu(img, 512, 4, "cuda", torch.bfloat16)[0]
for method in ("adain", "wavelet"):
    for name, fn in (("host", lambda: dc.finish(out, lq_pil, method, False, (128, 128), 4, None)),
                     ("GPU", lambda: dc.finish_gpu(out, lq_u8, method, False, (128, 128), 4, None))):
        for _ in range(3): fn()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(20): fn()
        torch.cuda.synchronize(); print(f"post-processing 512x512 {method}, {name}: {(time.perf_counter() - t0) / 20 * 1e3:.2f} ms/image")
PY
cat gpurun_out/summary.txt
