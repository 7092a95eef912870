"""Round 2: TMA-staged kernel, group size of the bin update (independent read-modify-write chains per lane).
   python profiles/scripts/r02_sweep_staged_ilp.py"""
import os, sys
sys.argv = sys.argv[:1]
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "r02_sweep_staged.py")).read()
exec(src.split("big = 1 << 26
for ns in (17, 24, 32, 48, 64):
    ref = None
    for ilp in (2, 1, 0, 4):
        t = run(ns, big, {"GEMHOOK_ACCT_STAGED": "1", "GEMHOOK_ACCT_STAGED_ILP": str(ilp)}, check=ref)
        ref = ref or t
for w, st in ((4, 4), (5, 2), (3, 6)):
    for ilp in (1, 0):
        run(64, big, {"GEMHOOK_ACCT_STAGED": "1", "GEMHOOK_ACCT_STAGED_ILP": str(ilp), "GEMHOOK_ACCT_WARPS": str(w), "GEMHOOK_ACCT_STAGES": str(st)})
for ilp in (1, 0):
    run(16, big, {"GEMHOOK_ACCT_STAGED": "1", "GEMHOOK_ACCT_STAGED_ILP": str(ilp)})
    run(8, big, {"GEMHOOK_ACCT_STAGED": "1", "GEMHOOK_ACCT_STAGED_ILP": str(ilp)})
