#!/bin/bash
# SASS evidence for profiles/: per kernel of liblinemod_b200.so the target architecture, register count and the
# instructions that show how it moves data (bulk async copies UBLKCP + mbarrier SYNCS, 128-bit loads, warp shuffles,
# REDUX, wide multiply-adds), with counts.  Runs without a GPU:  bash tools/sass_excerpt.sh > profiles/sass_rNN.txt
SO=${1:-6dpose_b200/csrc/liblinemod_b200.so}
echo "# cuobjdump -sass $SO   ($(date -u +%Y-%m-%dT%H:%MZ))"
cuobjdump -sass "$SO" | awk '
  /Fatbin elf code/ {next}
  /arch = / {arch=$3}
  /Function : / { if (fn != "") flush(); fn=$3; n=0; delete c; next }
  /^[[:space:]]+\/\*[0-9a-f]{4}\*\// {
     line=$0; sub(/^[[:space:]]+\/\*[0-9a-f]+\*\/[[:space:]]+/, "", line); sub(/^@!?U?P[0-9T]+[[:space:]]+/, "", line);
     split(line, t, /[ ;]/); op=t[1]; n++;
     if (op ~ /^(UBLKCP|SYNCS|LDG\.E\.128|LDG\.E\.64|LDG\.E\.CONSTANT|LDG\.E\.128\.CONSTANT|LDS|STS|STG|SHFL|REDUX|VOTE|ATOMG|RED|ATOMS|BAR|IMAD\.WIDE|LOP3|SHF|PRMT|MATCH|ACQBULK|UTMA|FENCE|MEMBAR|CCTL|ERRBAR|NANOSLEEP)/) { split(op, b, "."); key=b[1]; if (op ~ /\.128/) key=key ".128"; if (op ~ /\.64/ && key !~ /128/) key=key ".64"; c[key]++ }
  }
  function flush(   k, s) { s=""; for (k in c) s = s sprintf(" %s=%d", k, c[k]); printf "%-60s %5d instr |%s\n", fn, n, s }
  END { if (fn != "") flush(); print "# arch:", arch }'
echo
echo "# bulk async copy + mbarrier in k_coarse_packed<true> (TMA-staged bit-planes):"
cuobjdump -sass "$SO" | awk '/Function : .*k_coarse_packedILb1E/{f=1} /Function : /{if(!/k_coarse_packedILb1E/)f=0} f' | grep -E "UBLKCP|SYNCS" | sed -E 's/^\s+//' | head -12
echo
echo "# inner loop of k_refine_filter_w (shuffle -> wide multiply-add address -> two plane loads -> wrap shifts -> byte permute -> carry-save adders):"
cuobjdump -sass "$SO" | awk '/Function : .*k_refine_filter_w/{f=1} /Function : /{if(!/k_refine_filter_w/)f=0} f' | grep -E "SHFL.IDX|IMAD.WIDE.U32|LDG.E.CONSTANT|SHF.R.W|PRMT|LOP3.LUT" | sed -E 's/^\s+//' | sed -n '20,60p'
