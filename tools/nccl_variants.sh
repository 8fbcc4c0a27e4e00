run() { name=$1; shift; env "$@" timeout 120 python bench.py --gpus 8 --no-cfg5 --no-ess --no-cpu-baseline --no-other-methods --no-live-peaks --steps 30 --warmup 5 2> gpurun_out/r02_nccl_$name.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$name', round(d['ms_per_step'],4), d['step_breakdown_ms'], round(d['e2e']['ms_per_step'],4))"; }
run default X=1
run nvls NCCL_ALGO=NVLS
run tree NCCL_ALGO=Tree
run ring_ll128 NCCL_ALGO=Ring NCCL_PROTO=LL128
run ring_ll NCCL_ALGO=Ring NCCL_PROTO=LL
run nvlstree NCCL_ALGO=NVLSTree
run collnet_off NCCL_NVLS_ENABLE=0
run info NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL,TUNING
grep -i "nvls\|algo\|AllReduce" gpurun_out/r02_nccl_info.err | head -20
