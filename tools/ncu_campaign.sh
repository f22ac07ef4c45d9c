set -x
N="ncu --clock-control none --profile-from-start off"
timeout 300 $N --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2_v8_launches.csv python tools/ncu_targets.py --what step > gpurun_out/r2_v8_ncu_a.log 2>&1; echo "launch list rc=$?"
timeout 300 $N --set full --import-source on -k regex:conv_h3 -c 3 -o gpurun_out/r2_v8_conv_h3 python tools/ncu_targets.py --what enc > gpurun_out/r2_v8_ncu_b.log 2>&1; echo "h3 rc=$?"
timeout 300 $N --set full --import-source on -k regex:conv_ws -c 3 -o gpurun_out/r2_v8_conv_ws python tools/ncu_targets.py --what trunk > gpurun_out/r2_v8_ncu_c.log 2>&1; echo "ws rc=$?"
timeout 300 $N --set full -o gpurun_out/r2_v8_sif python tools/ncu_targets.py --what sif > gpurun_out/r2_v8_ncu_d.log 2>&1; echo "sif rc=$?"
timeout 400 $N --set full -o gpurun_out/r2_v8_small python tools/ncu_targets.py --what sinet,probclass,quant > gpurun_out/r2_v8_ncu_e.log 2>&1; echo "small rc=$?"
ls -la gpurun_out/ | head -20; du -sh gpurun_out
