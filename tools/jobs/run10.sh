python tools/jobs/dbg_tma.py 2>&1 | tail -5
