#!/bin/bash
# One gpurun call's worth of work; phases selected by arguments (tests bench ncu_bench ...).  Output -> gpurun_out/.
mkdir -p gpurun_out
for phase in "$@"; do
  case $phase in
    tests)   timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/tests.log; tail -3 gpurun_out/tests.log ;;
    bench)   timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 600 gpurun_out/bench_n1.err; head -c 300 gpurun_out/bench_n1.json; echo ;;
    ref)     timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2>&1 ;;
    pcie)    timeout 300 python tools/pcie_probe.py > gpurun_out/pcie_probe.txt 2>&1 ;;
    sweep)   timeout 900 python tools/codec_sweep.py > gpurun_out/codec_sweep.txt 2>&1; tail -5 gpurun_out/codec_sweep.txt ;;
    sweep_dq) timeout 900 python tools/codec_sweep.py 58720256 dequant > gpurun_out/codec_sweep_dq.txt 2>&1; timeout 900 python tools/codec_sweep.py 16777216 dequant > gpurun_out/codec_sweep_dq_16m.txt 2>&1; timeout 900 python tools/codec_sweep.py 4194304 dequant > gpurun_out/codec_sweep_dq_4m.txt 2>&1; tail -3 gpurun_out/codec_sweep_dq_4m.txt ;;
    iprobe)  timeout 120 tools/issue_probe > gpurun_out/issue_probe.txt 2>&1; cat gpurun_out/issue_probe.txt ;;
    kq)      timeout 600 tools/kq_sweep f16 > gpurun_out/kq_sweep_f16.txt 2>&1; timeout 300 tools/kq_sweep f32 > gpurun_out/kq_sweep_f32.txt 2>&1; cat gpurun_out/kq_sweep_f16.txt gpurun_out/kq_sweep_f32.txt ;;
    kq_ncu)  for t in T_Q4K T_Q6K; do timeout 600 ncu --set full --import-source on --clock-control none -k regex:quant_k_kernel -c 1 -f -o gpurun_out/kq_${t} tools/kq_sweep f16 $t first > gpurun_out/kq_ncu_${t}.log 2>&1; tail -2 gpurun_out/kq_ncu_${t}.log; done ;;
    ncu_batch) timeout 900 ncu --set full --import-source on --clock-control none -k regex:dequant_batch -c 1 -f -o gpurun_out/r02_dequant_batch python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_batch.log 2>&1; tail -2 gpurun_out/ncu_batch.log ;;
    launches) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launches.log 2>&1; tail -2 gpurun_out/ncu_launches.log; wc -l gpurun_out/r02_bench_launches.csv ;;
    c70_small) timeout 900 python tools/convert_llama70b.py 4 0 > gpurun_out/c70_small.log 2>&1; tail -4 gpurun_out/c70_small.log | cut -c1-1500 ;;
    c70_small2) GGQ_70B_SINGLE_RUNS=0 GGQ_70B_SHARDS=4G,1G GGQ_70B_ALSO_1GPU=0 timeout 900 python tools/convert_llama70b.py 6 0 > gpurun_out/c70_small.log 2>&1; tail -4 gpurun_out/c70_small.log | cut -c1-900 ;;
    c70_shards) GGQ_70B_SINGLE_RUNS=0 GGQ_70B_SHARDS=4G,2G,1G,512M GGQ_70B_ALSO_1GPU=0 timeout 1500 python tools/convert_llama70b.py 80 0 > gpurun_out/c70_shards.log 2>&1; grep "convert, -s" gpurun_out/c70_shards.log | cut -c1-400; tail -1 gpurun_out/c70_shards.log | cut -c1-600 ;;
    c70)     timeout 1700 python tools/convert_llama70b.py 80 0 > gpurun_out/c70.log 2>&1; tail -3 gpurun_out/c70.log | cut -c1-3000 ;;
    bench8)  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; tail -c 600 gpurun_out/bench_n8.err; head -c 400 gpurun_out/bench_n8.json; echo ;;
    bench4)  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; tail -c 600 gpurun_out/bench_n4.err; head -c 400 gpurun_out/bench_n4.json; echo ;;
    ncu_q40) timeout 600 ncu --set full --import-source on --clock-control none -k regex:quant_rows_oneshot -c 1 -f -o gpurun_out/r02_quant_q40_f16 python tools/codec_sweep.py 58720256 Q4_0:quant > gpurun_out/ncu_q40.log 2>&1; tail -2 gpurun_out/ncu_q40.log ;;
    ncu_kq)  for t in Q4K Q6K Q5K Q2K Q3K; do timeout 600 ncu --set full --import-source on --clock-control none -k regex:quant_k_kernel -c 1 -f -o gpurun_out/r02_quant_k_$t python tools/codec_sweep.py 58720256 $t:quant > gpurun_out/ncu_kq_$t.log 2>&1; tail -1 gpurun_out/ncu_kq_$t.log; done ;;
    ncu_kq2) for t in ${KQ_TYPES:-Q4K Q6K Q5K Q2K}; do timeout 400 ncu --set full --import-source on --clock-control none -k regex:quant_k_kernel -c 1 -f -o gpurun_out/r02b_quant_k_$t python tools/codec_sweep.py 58720256 $t:quant > gpurun_out/ncu_kq_$t.log 2>&1; tail -1 gpurun_out/ncu_kq_$t.log; done ;;
    dio)     df -h /var/tmp /tmp | cat; timeout 900 python tools/direct_io_probe.py > gpurun_out/dio.log 2>&1; tail -8 gpurun_out/dio.log ;;
    tests_conv) timeout 900 python -m pytest tests/test_convert.py -m gpu -x -q -s 2>&1 | tail -40 | cut -c1-400 ;;
    ncu_q8k) timeout 600 ncu --set full --import-source on --clock-control none -k regex:quant_rows_oneshot -c 1 -f -o gpurun_out/r02_quant_q8k_f16 python tools/codec_sweep.py 58720256 Q8K:quant > gpurun_out/ncu_q8k.log 2>&1; tail -2 gpurun_out/ncu_q8k.log ;;
    smoke)   timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ;;
    dq6)     timeout 600 tools/dq_sweep 40 58720256 > gpurun_out/dq_sweep_mode3_exact.txt 2>&1; grep -E "Q4K|Q5K|Q3K|Q2K|Q4_K|Q5_K" gpurun_out/dq_sweep_mode3_exact.txt | head -60 ;;
    small)   timeout 600 python tools/sanitize_small.py 2>&1 | tail -3 ;;
    absweep) # A/B libraries built by tools/build_ab.sh: AB_VARIANTS="default name1 name2", AB_ONLY=codec_sweep filter, AB_OUT=file
             timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "quantize or ragged or big_tensor_legacy or kat or fuzz or alignment" 2>&1 | tail -3
             for r in 1 2; do for v in $AB_VARIANTS; do
               so=$PWD/gguf_b200/libggq_ab_$v.so; [ $v = default ] && so=$PWD/gguf_b200/libggq.so
               GGQ_SO=$so timeout 300 python tools/codec_sweep.py 58720256 $AB_ONLY 2>&1 | grep -v " F32 " | sed "s/^/$v /"; done; done > gpurun_out/${AB_OUT:-absweep.txt}
             cat gpurun_out/${AB_OUT:-absweep.txt} ;;
    cpuprobe) timeout 900 python tools/cpu_port_probe.py > gpurun_out/cpu_port_probe.txt 2>&1; cat gpurun_out/cpu_port_probe.txt ;;
    soak)    timeout 900 python tools/soak.py ${SOAK_SECONDS:-300} ${SOAK_SEED:-1} > gpurun_out/soak.txt 2>&1; tail -14 gpurun_out/soak.txt ;;
    box)     bash tools/box_probe.sh > gpurun_out/box_probe.txt 2>&1 ;;
    bench2)  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -c 800 gpurun_out/bench_n2.err; head -c 300 gpurun_out/bench_n2.json; echo ;;
    mixtral2) timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/mixtral_sweep.py > gpurun_out/mixtral_n2.txt 2>&1; tail -13 gpurun_out/mixtral_n2.txt ;;
    tests_dq) timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "dequantize or slices or kat or nan_rule or alignment or big" 2>&1 | tail -4 ;;
    tests2)  timeout 1500 python -m pytest tests -m gpu -x -q -k "shard or slices or convert" 2>&1 | tail -5 ;;
    *) echo "unknown phase $phase" ;;
  esac
done
