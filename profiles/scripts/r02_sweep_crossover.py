"""Round 2: where the register-staged and the TMA-staged kernel cross (20..26 slots), and ring depth at the low end.
   python profiles/scripts/r02_sweep_crossover.py"""
import os
import sys
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "r02_sweep_staged.py")).read()
exec(src[:src.index("big = 1 << 26")].replace('"GEMHOOK_ACCT_STAGES")', '"GEMHOOK_ACCT_STAGES", "GEMHOOK_ACCT_STAGED_COLS")'))
big = 1 << 26
for ns in (18, 20, 21, 22, 23, 24, 26):
    ref = run(ns, big, {"GEMHOOK_ACCT_STAGED": "0"})
    run(ns, big, {"GEMHOOK_ACCT_STAGED": "1"}, check=ref)
for ns in (20, 22):
    for st in (2, 3, 6):
        run(ns, big, {"GEMHOOK_ACCT_STAGED": "1", "GEMHOOK_ACCT_STAGES": str(st)})
