#!/bin/bash
# SASS of the dominant kernels of the built library -> profiles/r2_sass_<name>.txt (instruction lines + opcode histogram)
lib=${1:-clair_torch_b200/lib/libclair_b200.so}
dump() {   # <mangled function> <out name> <demangled label>
  out=profiles/r2_sass_$2.txt
  { echo "# $3"; echo "# cuobjdump -sass -fun $1 (sm_100a; built by make -C clair_torch_b200/csrc)";
    cuobjdump -res-usage $lib 2>/dev/null | grep -A1 "$1" | tail -1 | sed 's/^/# /';
    cuobjdump -sass -fun "$1" $lib 2>/dev/null | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+//; s/\s*\/\* 0x[0-9a-f]+ \*\/$//' > /tmp/sass_body.txt
    echo "# opcode histogram (static):"; awk '{op=$2; if (op ~ /^@/) op=$3; sub(/\..*/, "", op); sub(/;/, "", op); n[op]++} END{for (k in n) printf "#   %-10s %d\n", k, n[k]}' /tmp/sass_body.txt | sort -k3 -n -r | head -24
    cat /tmp/sass_body.txt; } > $out
  echo "$out: $(grep -vc '^#' $out) instructions"
}
dump _ZN5clair22hdr_merge_fixed_kernelILi2ELi9ELi1ELb1ELi0EEEvNS_9HdrParamsE hdr_merge_c4 "clair::hdr_merge_fixed_kernel<2, 9, 1, true, 0> -- the headline c4 merge (2 px/thread, 9 frames in registers, fp32 val+std, single batch)"
dump _ZN5clair18pair_stats2_kernelILi2ELb1ELb1ELb1ELi1ELb0ELi128EEEvNS_10PairParamsE pair_stats_c3 "clair::pair_stats2_kernel<2, true, true, true, 1, false, 128> -- c3 linearity statistics (2 pair slots per warp, ERR, RELATIVE, FULL)"
dump _ZN5clair17pair_grad2_kernelILb0ELb1ELb0ELb0EEEvNS_10PairParamsE pair_grad_c2 "clair::pair_grad2_kernel<false, true, false, false> -- c2 table gradient (relative loss, no uncertainty weights)"
dump _ZN5clair17pair_grad2_kernelILb0ELb1ELb0ELb1EEEvNS_10PairParamsE pair_fused_c5 "clair::pair_grad2_kernel<false, true, false, true> -- c5 single-pair fused statistics + gradient pass"
dump _ZN5clair27hdr_merge_dark_strip_kernelILi5ELb1EEEvNS_9HdrParamsE dark_strip_c1 "clair::hdr_merge_dark_strip_kernel<5, true> -- c1 merge with the dark-field mix fused into its load, column-strip walk"
