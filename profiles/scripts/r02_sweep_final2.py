"""Round 2, final kernel set: register-staged (<= 22 slots), TMA-staged with 32 columns (23-37), with 16 columns (38-64) as
the host selects them, by slot count.  One JSON object per line.   python profiles/scripts/r02_sweep_final2.py"""
import os
import sys
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "r02_sweep_staged.py")).read()
exec(src[:src.index("big = 1 << 26")].replace('"GEMHOOK_ACCT_STAGES")', '"GEMHOOK_ACCT_STAGES", "GEMHOOK_ACCT_STAGED_COLS")'))
big = 1 << 26
for ns in (1, 2, 4, 8, 12, 16, 17, 20, 24, 32, 38, 39, 40, 48, 56, 64):
    run(ns, big)
for ns in (17, 20, 24):
    run(ns, big, {"GEMHOOK_ACCT_STAGED": "0"})
for ns in (32, 38, 40, 48):
    run(ns, big, {"GEMHOOK_ACCT_STAGED_COLS": "16"})
    run(ns, big, {"GEMHOOK_ACCT_STAGED_COLS": "32"})
