#!/bin/bash
# 2-GPU box: multi-GPU tests, then bench at N = 1 and N = 2
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo2.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 800 > gpurun_out/r2_pytest14_multi.log 2>&1
tail -3 gpurun_out/r2_pytest14_multi.log
timeout 600 python bench.py --gpus 1 > gpurun_out/r2_bench14_n1.json 2> gpurun_out/r2_bench14_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r2_bench14_n2.json 2> gpurun_out/r2_bench14_n2.err
python - <<'PY'
import json
for n in (1,2):
    try:
        d=json.loads(open(f"gpurun_out/r2_bench14_n{n}.json").read().strip().splitlines()[-1])
        print(n, "value %.1f G ms %.3f e2e %.1f G (%.3f ms) kernel %.3f" % (d["value"]/1e9, d["ms_per_step"], d["e2e"]["value"]/1e9, d["e2e"]["ms_per_step"], d["roofline"]["kernel_ms_per_step"]), d["config"].get("cpu_affinity"), d["parity"].get("vs_single_gpu",{}).get("bitwise_equal"))
        for k,v in d.get("other_workloads",{}).items(): print("   ", k, "value %.1f G e2e %.1f G" % (v["value"]/1e9, v["e2e"]["value"]/1e9) if "value" in v else v)
    except Exception as ex: print(n, "failed", ex)
PY
tail -3 gpurun_out/r2_bench14_n2.err
