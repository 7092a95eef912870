"""Round 2: 2 KB ring buffers and more than eight warps per SM (16 columns) at the fullest slot tables -- the variant this script
   drove (GEMHOOK_ACCT_STAGE_ROWS=4) was removed after the measurement: profiles/r02_acct_staged_variants.jsonl, last 21 rows."""
import os
import sys
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "r02_sweep_staged.py")).read()
exec(src[:src.index("big = 1 << 26")].replace('"GEMHOOK_ACCT_STAGES")', '"GEMHOOK_ACCT_STAGES", "GEMHOOK_ACCT_STAGED_COLS", "GEMHOOK_ACCT_STAGE_ROWS")'))
big = 1 << 26
for ns in (64, 56, 48):
    ref = run(ns, big)
    for w, st in ((8, 3), (9, 3), (10, 2), (8, 2)):
        run(ns, big, {"GEMHOOK_ACCT_STAGE_ROWS": "4", "GEMHOOK_ACCT_WARPS": str(w), "GEMHOOK_ACCT_STAGES": str(st)}, check=ref)
for n in (513, 4097, (1 << 20) + 77):
    ref = run(64, n, {"GEMHOOK_ACCT_SMALL": "0"}, reps=4)
    run(64, n, {"GEMHOOK_ACCT_SMALL": "0", "GEMHOOK_ACCT_STAGE_ROWS": "4", "GEMHOOK_ACCT_WARPS": "9"}, reps=4, check=ref)
