timeout -s INT 120 profiles/probes/gather_span > gpurun_out/gather_span.log 2>&1
cat gpurun_out/gather_span.log
export BKX_TRACE=1
timeout -s INT 240 python profiles/ab_kernel.py --reps 1 --prefix-k 16 --set "INPUT=packed2 BKX_WAVE=0" --set "INPUT=packed2 BKX_WAVE=1" > gpurun_out/wave_k16.log 2>&1
grep -v finish_index gpurun_out/wave_k16.log | tail -14
