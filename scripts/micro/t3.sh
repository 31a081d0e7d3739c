t() { timeout 120 python scripts/run_one.py --D 192 --reps 4 | tail -2 | head -1 | tr "," "\n" | grep scanline | grep -v "#n" | tr "\n" " "; echo; }
echo default; t
echo NWH=15; TSM_S3_NWH=15 t; TSM_S3_NWH=15 t
echo NWH=15 P4 N4; TSM_S3_NWH=15 TSM_S3_P=4 TSM_S3_NSTH=4 t
echo NWH=8; TSM_S3_NWH=8 t
echo NWH=5; TSM_S3_NWH=5 t
echo NWH=3; TSM_S3_NWH=3 t
