timeout 300 python profiles/tools/sa_b3_ab.py 2,5 $((16384+128)) 2>&1 | grep -v "^Trace" | grep -v "preload=0" | tail -8
