run() { label=$1; shift
  env "$@" timeout 250 python bench.py --steps 2 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); c=d['config']; print('$label', 'rtfx %.0f'%d['value'], 'dec_ms %.1f'%c['ms_decode'], 'tok %.0f'%c['tokens_per_step'], [e.get('avg_launch_us') for e in [d['roofline']]+d['roofline_extra'] if 'self' in e['kernel']])"
}
timeout 300 python -m pytest tests/test_decoder_gpu.py tests/test_host_gpu.py -x -q 2>&1 | tail -2
run small_sa A=1
run turbo_sa SB_BENCH_ARCH=large-v3-turbo
