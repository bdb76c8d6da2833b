#!/bin/bash
# SASS evidence of the Blackwell-native instructions in libtmf.so: per kernel, counts of the mnemonics B200_PROFILING.md lists
# (UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UBLKCP = TMA / bulk copies, UTCBAR = tcgen05.commit), the packed
# fp32x2 instructions of the training kernels (FADD2/FFMA2), warp-level REDUX, and the legacy tensor path (HMMA: must be 0).
SO=${1:-teamoflow_b200/csrc/libtmf.so}
echo "# $(basename $SO)  $(date -u +%F)  nvcc $(/usr/local/cuda/bin/nvcc --version | grep release | sed 's/.*release //')"
echo "# UTCHMMA UTCBAR  LDTM UTMALDG UTMAPF UBLKCP SYNCS FADD2 FFMA2 FMNMX3 REDUX HMMA | kernel"
/usr/local/cuda/bin/cuobjdump -sass $SO | awk '
/Function :/ {name=$3; next}
{ if (name=="") next;
  n[name]++;
  if ($0 ~ /UTCHMMA/) a[name]++; if ($0 ~ /UTCBAR/) b[name]++; if ($0 ~ /LDTM/) c[name]++; if ($0 ~ /UTMALDG/) d[name]++;
  if ($0 ~ /UTMAPF/) pf[name]++; if ($0 ~ /UBLKCP/) e[name]++; if ($0 ~ /SYNCS/) f[name]++; if ($0 ~ /FADD2/) g[name]++; if ($0 ~ /FFMA2/) h[name]++;
  if ($0 ~ /FMNMX3/) m3[name]++; if ($0 ~ /REDUX/) i[name]++; if ($0 ~ /HMMA/ && $0 !~ /UTCHMMA/) j[name]++; }
END { for (k in n) if (a[k]+c[k]+d[k]+e[k]+g[k]+h[k]+i[k]+j[k] > 0) printf "%6d %6d %5d %7d %6d %6d %5d %5d %5d %6d %5d %4d | %s\n", a[k], b[k], c[k], d[k], pf[k], e[k], f[k], g[k], h[k], m3[k], i[k], j[k], k }' | c++filt | sed 's/(CUtensorMap_st.*//; s/(tmf::[A-Za-z]*Params.*//; s/(float const\*.*//; s/(long long.*//; s/(int.*//' | sort -t'|' -k2
