python -c "from __graft_entry__ import smoke; smoke()" 2>&1 | tail -3
