#!/bin/bash
# which realization of the classic ensemble fails? groups of 8, then single seeds of the failing groups
for g in 16 24 32 40 48 56; do
  if timeout 300 python profiles/ens_soak.py 8 8 500 4 $g 2>&1 | grep -q "^OK"; then echo "group $g ok"; else
    echo "group $g FAILS"
    for k in 0 1 2 3 4 5 6 7; do
      s=$((g + k))
      if timeout 200 python profiles/ens_soak.py 1 1 500 4 $s 2>&1 | grep -q "^OK"; then :; else echo "  index $s (seed $((1000 + s))) FAILS"; fi
    done
  fi
done
