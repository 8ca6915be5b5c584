#!/bin/bash
# Strong scaling of the sharded run_MCMC on N GPUs of one box: fixed ensembles (C4 512 x 4096, C5 1024 x 1024, C3 256 x 500).
#   tools/strong_scaling.sh N [out.jsonl]
N=${1:-2}; OUT=${2:-gpurun_out/strong_${N}gpu.jsonl}
mkdir -p "$(dirname "$OUT")"; : > "$OUT"
for w in C4 C5 C3; do
  if [ "$N" -gt 1 ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29541 \
      bench.py --scaling strong --workload $w 2> "gpurun_out/strong_${N}gpu_${w}.err" | tail -1 >> "$OUT"
  else
    timeout 600 python bench.py --scaling strong --workload $w 2> "gpurun_out/strong_${N}gpu_${w}.err" | tail -1 >> "$OUT"
  fi
done
python - "$OUT" <<'PY'
import json, sys
for line in open(sys.argv[1]):
    try:
        d = json.loads(line)
    except Exception:
        print("unparsable:", line[:200]); continue
    print("%s on %d GPU(s): %.3f ms/iteration, %.0f evals/s, host ms %s, digest %s" % (d["config"]["workload"][:2], d["n_gpus"], d["ms_per_step"], d["value"],
          {k: round(v, 3) for k, v in d["host_ms_per_iteration"].items()}, d["chain_digest"]))
PY
