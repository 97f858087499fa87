#!/bin/bash
# N-GPU evidence run: bench at N, parity of both exchanges, split of the fused exchange with NVLink counters,
# the multi-GPU C ABI (NCCL form) and the CLI.   usage: run_gpu_multi.sh N
N=${1:-8}
mkdir -p gpurun_out
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
{
nvidia-smi -L | head -8
echo "== bench N=$N"
run 29551 bench.py --gpus $N 2>&1 | grep "^{" | tee gpurun_out/r2_bench_n$N.json | cut -c1-600
echo "== check_fused C N=$N"
run 29552 tools/check_fused_exchange.py C 2>&1 | grep "^{" | tee gpurun_out/r2_check_fused_n$N.json
echo "== fused parts + nvlink N=$N"
run 29553 tools/time_fused_parts.py 2>&1 | grep "^{" | tee gpurun_out/r2_fused_parts_n$N.json
for role in uniq depth; do
  echo "== FGFA_X_ROLE=$role"
  FGFA_X_ROLE=$role run 29554 tools/time_fused_parts.py 2>&1 | grep "^{" | tee gpurun_out/r2_fused_parts_n${N}_$role.json | cut -c1-400
done
echo "== single-process C ABI, peer exchange, NVLink bytes of kernel X (ncu)"
timeout 600 python tools/multi_peer_probe.py $N peer 5 2>&1 | grep "^{" | tee gpurun_out/r2_multi_peer_n$N.json
timeout 600 python tools/multi_peer_probe.py $N nccl 5 2>&1 | grep "^{" | tee gpurun_out/r2_multi_nccl_n$N.json
PROBE_VERIFY=0 timeout 900 ncu --metrics nvlrx__bytes.sum,nvltx__bytes.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_aperture_peer.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_uniq_exchange -c $((2*N)) --csv --log-file gpurun_out/r2_x_nvlink_n$N.csv python tools/multi_peer_probe.py $N peer 1 2>&1 | tail -2
tail -n +1 gpurun_out/r2_x_nvlink_n$N.csv | cut -c1-300 | tail -30
echo "== multi ABI + CLI"
timeout 600 python -m pytest tests/test_multi_gpu_abi.py -q -m gpu 2>&1 | tail -5
} > gpurun_out/r2_multi_n$N.log 2>&1
cat gpurun_out/r2_multi_n$N.log
