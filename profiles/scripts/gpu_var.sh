cd $GRAFT_REPO_ROOT
for V in "" "-DGEMHOOK_EXP_NO_PUBLISH"; do
  rm -rf kubeshare_b200/csrc/build kubeshare_b200/lib
  make -s -C kubeshare_b200/csrc VARIANT="$V" > /dev/null 2>&1
  python bench.py --only-roofline --steps 10 --warmup 3 2>/dev/null | grep "^{" | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('[$V]', 'big %.1f GB/s (min %.2f us)'%(d['ring_2p26']['gbps'], d['ring_2p26']['min_ms']*1e3), 'small avg %.1f us min %.1f us'%(d['ring_2p20']['avg_ms']*1e3, d['ring_2p20']['min_ms']*1e3), 'grid', d['ring_2p26']['grid'], d['ring_2p20']['grid'])" || echo "FAILED $V"
done
rm -rf kubeshare_b200/csrc/build kubeshare_b200/lib; make -s -C kubeshare_b200/csrc > /dev/null 2>&1
python -m pytest tests/test_gpu_acct.py -x -q -m gpu 2>&1 | tail -2
