timeout 280 python -c "
import faulthandler,sys
faulthandler.dump_traceback_later(200, exit=True)
sys.argv=['bench.py','--steps','20','--warmup','5']
import runpy
runpy.run_path('bench.py', run_name='__main__')
" > gpurun_out/dbg_bench.json 2> gpurun_out/dbg_bench.err
echo rc=$?; tail -30 gpurun_out/dbg_bench.err | cut -c1-200; head -c 300 gpurun_out/dbg_bench.json
