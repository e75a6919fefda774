# density exchange flavours at the per-rank size of the 8-GPU headline run:  bash tools/fused_probe.sh [gpus] [particles]
G=${1:-2}; P=${2:-$((125000000 * G))}
run() { tag=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $G --particles $P --warmup 5 --steps 60 --no-roundtrip --no-cpu --no-fp32 --no-batched --no-single "$@" > gpurun_out/fp_$tag.json 2> gpurun_out/fp_$tag.err; python -c "
import json
d=json.loads(open('gpurun_out/fp_$tag.json').read().strip().splitlines()[-1])
print('$tag', 'G/s', round(d['value']/1e9,1), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']/1e9,1), d['config'].get('collective'), 'mc', d['config'].get('fused_multicast'), d['config'].get('gather'), d['config']['parity']['rho_crc'], d['clocks'].get('sm_mhz'))
" || tail -5 gpurun_out/fp_$tag.err; }
run nccl --collective nccl
run fusedmc --collective fused
PIC_FUSED_MULTICAST=0 run fusedp2p --collective fused

run fusedmc2 --collective fused
