set -x
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
