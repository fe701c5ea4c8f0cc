#!/usr/bin/env bash
# Evidence that the kernels are Blackwell-native: count tensor-core / TMA / TMEM SASS mnemonics per kernel.
cd "$(dirname "$0")/.."
out=profiles/sass_summary.txt
: > $out
for o in lca_b200/ops/build/fmha_fwd_sm100.o lca_b200/ops/build/fmha_bwd_sm100.o lca_b200/ops/build/fmha_fwd_fp8_sm100.o; do
  for fn in $(cuobjdump -sass $o | grep -oE "Function : [^ ]+" | awk '{print $3}' | sort -u); do
    body=$(cuobjdump -sass -fun "$fn" $o)
    echo "== $(echo $fn | c++filt)" >> $out
    for m in "UTCHMMA" "UTCQMMA" "UTCBAR" "UTMALDG" "UBLKCP.S.G" "UBLKCP.G.S" "LDTM" "STTM" "UTCATOMSWS" "USETMAXREG" "SYNCS" "LDG.E.STRONG.SYS" "STG.E.STRONG.SYS" "REDG.E.ADD.STRONG.SYS" "MEMBAR.*SYS" "STG.E.128" "STG.E.128.STRONG.SYS" "[^C]HMMA" "MUFU.EX2" "FFMA2" "FADD2" "FMUL2"; do
      n=$(echo "$body" | grep -cE "$m")
      echo "   $m: $n" >> $out
    done
  done
done
echo "$(grep -c "^==" $out) kernels summarised in $out"
