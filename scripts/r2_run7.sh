set -x
python scripts/repro_wkt.py small all > gpurun_out/r2_repro_small.log 2>&1
python scripts/repro_wkt.py default all > gpurun_out/r2_repro_default.log 2>&1
timeout 600 compute-sanitizer --tool initcheck --track-unused-memory no python scripts/repro_wkt.py small one > gpurun_out/r2_repro_initcheck.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python scripts/repro_wkt.py small one > gpurun_out/r2_repro_memcheck.log 2>&1
grep -c "DIFF" gpurun_out/r2_repro_small.log gpurun_out/r2_repro_default.log
grep "DIFF\|items" gpurun_out/r2_repro_small.log | head -20
grep -i "error\|uninit" gpurun_out/r2_repro_initcheck.log | head -10
grep -i "error\|invalid" gpurun_out/r2_repro_memcheck.log | head -10
