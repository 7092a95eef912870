"""Round 2: TMA-staged kernel, group size of the bin update (independent read-modify-write chains per lane).
   python profiles/scripts/r02_sweep_staged_ilp.py"""
import os, sys
sys.argv = sys.argv[:1]
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "r02_sweep_staged.py")).read()
exec(src.split("big = 1 << 26")[0].replace('"GEMHOOK_ACCT_STAGES")', '"GEMHOOK_ACCT_STAGES", "GEMHOOK_ACCT_STAGED_ILP")'))
big = 1 << 26
for ns in (24, 32, 48, 64):
    ref = None
    for ilp in (2, 4, 8):
        t = run(ns, big, {"GEMHOOK_ACCT_STAGED": "1", "GEMHOOK_ACCT_STAGED_ILP": str(ilp)}, check=ref)
        ref = ref or t
for w, st in ((4, 4), (4, 5), (5, 2)):
    for ilp in (4, 8):
        run(64, big, {"GEMHOOK_ACCT_STAGED": "1", "GEMHOOK_ACCT_STAGED_ILP": str(ilp), "GEMHOOK_ACCT_WARPS": str(w), "GEMHOOK_ACCT_STAGES": str(st)})
for ilp in (4, 8):
    run(48, big, {"GEMHOOK_ACCT_STAGED": "1", "GEMHOOK_ACCT_STAGED_ILP": str(ilp), "GEMHOOK_ACCT_WARPS": "5", "GEMHOOK_ACCT_STAGES": "4"})
    run(16, big, {"GEMHOOK_ACCT_STAGED": "1", "GEMHOOK_ACCT_STAGED_ILP": str(ilp)})
