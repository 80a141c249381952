#!/bin/bash
# 8-GPU box: multi-GPU tests (2 / 4 / 8 ranks), then bench at N = 4 and N = 8
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo8.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 800 > gpurun_out/r2_pytest17_multi.log 2>&1
tail -3 gpurun_out/r2_pytest17_multi.log
for n in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n > gpurun_out/r2_bench17_n$n.json 2> gpurun_out/r2_bench17_n$n.err
done
python - <<'PY'
import json
for n in (4,8):
    try:
        d=json.loads(open(f"gpurun_out/r2_bench17_n{n}.json").read().strip().splitlines()[-1])
        print(n, "value %.1f G ms %.3f e2e %.1f G (%.3f ms) kernel %.3f" % (d["value"]/1e9, d["ms_per_step"], d["e2e"]["value"]/1e9, d["e2e"]["ms_per_step"], d["roofline"]["kernel_ms_per_step"]), d["config"].get("cpu_affinity"), d["parity"].get("vs_single_gpu",{}).get("bitwise_equal"), "h2d", d["e2e"]["h2d_bytes_per_step"])
        for k,v in d.get("other_workloads",{}).items(): print("   ", k, "value %.1f G e2e %.1f G" % (v["value"]/1e9, v["e2e"]["value"]/1e9) if "value" in v else v, v.get("parity",{}).keys() if isinstance(v,dict) else "")
    except Exception as ex: print(n, "failed", ex)
PY
tail -2 gpurun_out/r2_bench17_n8.err
