#!/bin/bash
# Counts of the Blackwell-specific SASS mnemonics in the built library (no GPU needed): the evidence that the hot
# contractions are tcgen05 (UTCHMMA) with TMEM accumulators (LDTM / STTM) fed by TMA (UTMALDG / UTMASTG / UTMAREDG).
#   bash scripts/sass_summary.sh > profiles/sass_summary.txt
LIB=${1:-iggt_official_b200/lib/libiggt_b200.so}
echo "# cuobjdump -sass $LIB  ($(date -u +%Y-%m-%d), $(nvcc --version | tail -2 | head -1))"
cuobjdump -sass "$LIB" > /tmp/iggt_sass.txt
echo "## library totals"
for m in UTCHMMA UTCHMMA.2CTA UTCQMMA LDTM STTM UTMALDG UTMASTG UTMAREDG UTMAPF UTCBAR SYNCS FFMA2 FADD2 FMUL2 MUFU.EX2 MUFU.RCP HMMA ; do
  printf "%-14s %6d\n" $m $(grep -c -- "$m" /tmp/iggt_sass.txt)
done
echo "## per kernel: UTCHMMA / LDTM / STTM / UTMALDG / UTMASTG|UTMAREDG / FFMA2+FADD2+FMUL2 / MUFU"
awk '/Function : /{name=$3} {if(name!=""){ if($0 ~ /UTCHMMA/) a[name]++; if($0 ~ /LDTM/) b[name]++; if($0 ~ /STTM/) c[name]++; if($0 ~ /UTMALDG/) d[name]++; if($0 ~ /UTMASTG|UTMAREDG/) e[name]++; if($0 ~ /FFMA2|FADD2|FMUL2/) f[name]++; if($0 ~ /MUFU/) g[name]++; seen[name]=1}} END{for(n in seen) if(a[n]+b[n]+d[n]>0) printf "%s %d %d %d %d %d %d %d\n", n, a[n],b[n],c[n],d[n],e[n],f[n],g[n]}' /tmp/iggt_sass.txt | sort | while read n a b c d e f g; do printf "%-110s %4d %4d %4d %4d %4d %5d %4d\n" "$(echo $n | c++filt | cut -c1-110)" $a $b $c $d $e $f $g; done
