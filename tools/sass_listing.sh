#!/bin/bash
# SASS of one kernel of the shipped library + opcode histogram + the lines that prove the TMA / mbarrier path.
#   tools/sass_listing.sh 'k_stripILi2ELi3ELi9E' > profiles/r2_sass_k_strip.txt
set -e
LIB=wobble_jax_b200/libjabble_b200.so
PAT=${1:-k_stripILi2ELi3ELi9E}
TMP=$(mktemp)
cuobjdump -sass "$LIB" | awk -v pat="$PAT" '/Function : /{f = index($0, pat) > 0} f' > "$TMP"
echo "# cuobjdump -sass $LIB, function matching $PAT ($(grep -c '^\s*/\*[0-9a-f]\{4,\}\*/' "$TMP") instructions)"
echo "# library source hash $(cat wobble_jax_b200/libjabble_b200.so.hash)"
echo "## opcode histogram (static)"
grep -oE "^\s+/\*[0-9a-f]+\*/\s+(@!?U?P[0-9T]+ )?[A-Z0-9_.]+" "$TMP" | awk '{print $NF}' | sed 's/\..*//' | sort | uniq -c | sort -rn | head -40
echo "## bulk async copy (TMA engine) and mbarrier instructions"
grep -nE "UBLKCP|SYNCS|UTMA" "$TMP" | sed 's/  */ /g' | cut -c1-140
echo "## listing"
grep -E "^\s+/\*[0-9a-f]{4,}\*/" "$TMP" | sed 's#/\* 0x[0-9a-f]* \*/##' | sed 's/  *$//'
rm -f "$TMP"
