#!/usr/bin/env bash
# round-2 call J (2 GPUs): GC / StereoNet on tcgen05 vs reference golden, 2-GPU tests, bench.py under torchrun (train block + NCCL)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -s -k "gc_and_stereonet" --timeout 300 > gpurun_out/pytest_j1.log 2>&1; echo "gc/stereonet tests=$?"; grep -E "max \|d\||passed|failed" gpurun_out/pytest_j1.log | tail -8
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_autograd_paths.py -m gpu -q -k "two_gpus or tensor_device" --timeout 600 > gpurun_out/pytest_j2.log 2>&1; echo "2-gpu tests=$?"; tail -4 gpurun_out/pytest_j2.log
NCCL_DEBUG=INFO timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_j_n2.json 2> gpurun_out/bench_j_n2.err; echo "bench n2=$?"
wc -l gpurun_out/bench_j_n2.json; grep -c "NCCL INFO" gpurun_out/bench_j_n2.err; grep -E "nRanks|NVLS" gpurun_out/bench_j_n2.err | head -4
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_j_n2.json"))
print("N=2 pairs/s %.1f ms/step %.2f e2e %.1f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
print("train", json.dumps(d.get("train"))[:900])
PY
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_j_n1.json 2> gpurun_out/bench_j_n1.err; echo "bench n1=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_j_n1.json"))
print("N=1 pairs/s %.1f ms/step %.2f e2e %.1f segments %s frac %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], {k: round(v, 3) for k, v in d["segments_ms"].items()}, d["roofline"]["frac"]))
print("train", d["train"]["ms_per_step"], d["train"]["pairs_per_s"])
print("ops", json.dumps({k.split(" ")[0]: (v["ms"], v["frac_of_hbm_peak"]) for k, v in d["ops"].items()}))
print("gpu_torch", json.dumps(d["gpu_torch_baseline"])[:400])
PY
