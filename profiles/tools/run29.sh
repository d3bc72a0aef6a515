timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout -s KILL 400 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | cut -c1-600
