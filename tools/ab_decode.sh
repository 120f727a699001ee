run() { # label, env...
  label=$1; shift
  env "$@" timeout 200 python bench.py --steps 2 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); c=d['config']; print('$label', 'rtfx %.0f'%d['value'], 'dec_ms %.1f'%c['ms_decode'], 'steps %.0f'%c['decoder_steps_per_step'], 'us/step %.0f'%(1000*c['ms_decode']/c['decoder_steps_per_step']), 'tok %.0f'%c['tokens_per_step'], 'launches', d['gpu_launches'])"
}
timeout 300 python -m pytest tests/test_decoder_gpu.py -x -q 2>&1 | tail -3
run split4_l2 A=1
run split2_l2 SB_DEC_FC2_SPLITS=2
run split1_l2 SB_DEC_FC2_SPLITS=1
run split4_l1 SB_DECODE_LANES=1
run split4_l3 SB_DECODE_LANES=3
