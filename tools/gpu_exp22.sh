#!/bin/bash
# experiment 22: full GPU suite after the fixture rename
cd /root/repo
echo "== parity"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
