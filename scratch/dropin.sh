#!/bin/bash
# usage: dropin.sh <case>
set -e
cd /root/repo
D=$(mktemp -d)
python - <<PY
import sys; sys.path.insert(0,'tests'); import parity_common as pc
open('$D/m.inp','w').write(pc.case_inp('$1'))
PY
oracle/_ref/runswmm $D/m.inp $D/ref.rpt $D/ref.out > /dev/null
LD_PRELOAD=$PWD/tests/emul/libswmm5_b200_seam_emul.so oracle/_ref/runswmm $D/m.inp $D/b200.rpt $D/b200.out > $D/b200.log 2>&1 || (tail -5 $D/b200.log; grep -i error $D/b200.rpt | head)
cmp $D/ref.out $D/b200.out && echo "$1: .out IDENTICAL ($(stat -c %s $D/ref.out) bytes)"
diff <(grep -v "Analysis begun\|Analysis ended\|Total elapsed" $D/ref.rpt) <(grep -v "Analysis begun\|Analysis ended\|Total elapsed" $D/b200.rpt) > $D/rpt.diff && echo "$1: .rpt IDENTICAL" || (echo "$1: .rpt differs:"; head -20 $D/rpt.diff)
