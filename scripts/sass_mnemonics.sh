#!/bin/bash
# Counts of the SASS mnemonics that show tcgen05 / TMEM / bulk-copy (TMA engine) / mbarrier / setmaxnreg use, per kernel
# of libmsmp_b200.so (cuobjdump -sass on the in-tree objects).  Output: markdown table on stdout.
cd "$(dirname "$0")/../msmp_pde_b200/csrc/build"
echo "| kernel | UTCHMMA (tcgen05.mma) | LDTM (tcgen05.ld) | STTM (tcgen05.st) | UBLKCP (cp.async.bulk) | UTMALDG (cp.async.bulk.tensor) | SYNCS (mbarrier) | USETMAXREG | UTCBAR (tcgen05.commit) |"
echo "|---|---:|---:|---:|---:|---:|---:|---:|---:|"
for o in linear_ts linear_tma linear_tc wgrad_ws wgrad_tc edge_tc edge_ws lem_tc; do
  cuobjdump -sass $o.o | awk -v obj=$o '
    /Function :/ { if (name != "") print_row(); name=$3; for (k in c) delete c[k]; next }
    { for (i = 1; i <= NF; ++i) { m=$i; sub(/\..*/, "", m);
        if (m=="UTCHMMA"||m=="LDTM"||m=="STTM"||m=="UBLKCP"||m=="UTMALDG"||m=="SYNCS"||m=="USETMAXREG"||m=="UTCBAR") c[m]++ } }
    function print_row() { if (c["UTCHMMA"] > 0) printf("| `%s` (%s.cu) | %d | %d | %d | %d | %d | %d | %d | %d |\n", name, obj, c["UTCHMMA"], c["LDTM"], c["STTM"], c["UBLKCP"], c["UTMALDG"], c["SYNCS"], c["USETMAXREG"], c["UTCBAR"]) }
    END { print_row() }'
done | sed 's/_ZN4msmp[0-9]*//; s/EvNS_[0-9A-Za-z_]*E`/`/'
