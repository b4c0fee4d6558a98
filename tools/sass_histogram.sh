#!/bin/bash
# SASS evidence: per-kernel opcode histogram of the built library (run in the build container; needs no GPU).
#   bash tools/sass_histogram.sh > profiles/r02_sass_opcodes.txt
cd "$(dirname "$0")/.."
LIB=co-detr-tensorrt_b200/csrc/libmsda_b200.so
echo "# cuobjdump -sass $LIB  ($(date -u +%F), $(nvcc --version | tail -2 | head -1))"
echo "# whole library: the mnemonics that identify Blackwell / TMA / tcgen05 code paths"
cuobjdump -sass $LIB | grep -oE "\b(FHFMA|FFMA2|HFMA2|LDG\.E\.128\.CONSTANT|LDG\.E\.ENL2\.256[A-Z.]*|LDS\.128|UBLKCP[A-Z0-9.]*|UBLKPF[A-Z0-9.]*|UTMALDG[A-Z0-9.]*|UTMASTG[A-Z0-9.]*|UTCHMMA[A-Z0-9.]*|UTCBAR[A-Z0-9.]*|LDTM[A-Z0-9.x]*|SYNCS[A-Z0-9.]*|REDG\.E\.ADD\.[A-Z0-9x.]*|SHFL\.(IDX|BFLY|UP)|ACQBULK|CCTL[A-Z.]*)" | sort | uniq -c | sort -rn
for pat in 'msda_fwd_hpI6__halfLi1ELi8ELb0ELb0ELb0' 'msda_fwd_hpI13__nv_bfloat16Li2ELi8ELb0ELb0ELb0' 'msda_fwd_hpIfLi0ELi8ELb0ELb0ELb0' 'msda_fwd_vecI6__halfLi32ELi4ELi1ELi1ELb0ELb0ELb0' 'msda_fwd_smallI6__halfLi32ELi1' 'value_proj_persistent_kernelILi0ELi0'; do
  echo
  echo "## kernel matching $pat"
  cuobjdump -sass $LIB | awk -v pat="$pat" '/Function : /{f=0} $0 ~ ("Function : .*" pat) {f=1; print "# " $3} f' | grep -v "^\s*/\* 0x" | grep -E "^\s+/\*[0-9a-f]{4}\*/" \
    | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//; s/^@!?U?P[0-9T]+\s+//' | awk '{print $1}' | sed -E 's/;$//' | sort | uniq -c | sort -rn | head -28
done
