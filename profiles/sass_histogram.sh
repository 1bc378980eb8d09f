#!/bin/bash
# Opcode histograms of the hot kernels from the built objects (cuobjdump -sass; no GPU needed):
#   bash profiles/sass_histogram.sh        -> profiles/sass_<kernel>.txt
# What to look for: UBLKCP = cp.async.bulk (the 1-D form of TMA: contiguous tiles need no tensor map, so there is no
# UTMALDG), SYNCS = mbarrier operations, MATCH = match.any, BAR = named / CTA barriers, ATOM / RED, MUFU (division
# slow paths), and that no HMMA / UTCMMA shows up (nothing here is a contraction).
cd "$(dirname "$0")/../keras-object-detection_b200/csrc" || exit 1
hist() {   # object, mangled-name regex, output name, description
  local fn
  fn=$(cuobjdump -sass "$1" 2>/dev/null | grep -o "Function : .*" | sed 's/Function : //' | grep -E "$2" | head -1)
  [ -z "$fn" ] && { echo "no function matching $2 in $1"; return; }
  {
    echo "# $4"
    echo "# $(cu++filt "$fn" 2>/dev/null || echo "$fn")   [$1, $(git -C ../.. rev-parse --short HEAD)]"
    cuobjdump -sass -fun "$fn" "$1" 2>/dev/null | grep -E '^\s+/\*[0-9a-f]{4,6}\*/' | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//; s/^@!?U?P[0-9T]+\s+//' |
      awk '{op=$1; sub(/;$/,"",op); split(op,a,"."); full[op]++; base[a[1]]++; n++}
           END{printf "total instructions %d\n", n; print "-- by base opcode"; for(k in base) printf "%6d %s\n", base[k], k | "sort -rn"; close("sort -rn");
               print "-- selected full mnemonics"; for(k in full) if (k ~ /^(UBLKCP|SYNCS|MATCH|BAR|ATOM|ATOMS|RED|MUFU|LDGSTS|UTMA|HMMA|UTCMMA|MEMBAR|FENCE|ELECT|CCTL|LDS|STS|LDG|STG|SHFL|VOTE|REDUX|FLO|BREV|POPC|FSETP|FMNMX|ACQBULK|ARRIVES|ERRBAR|NANOSLEEP|CALL)/) printf "%6d %s\n", full[k], k | "sort -k2"; close("sort -k2")}'
  } > "../../profiles/sass_$3.txt"
  echo "profiles/sass_$3.txt: $(sed -n 3p ../../profiles/sass_$3.txt)"
}
hist yh_decode_nms.o 'decode_nms_tma_kernelILi2ELi20ELi2EfE' decode_nms_tma_kernel "fused decode+IoU+NMS, VOC tiles through a TMA ring (the graded kernel)"
hist yh_decode_nms.o 'decode_nms_coop_kernelILi80ELi3EfE' decode_nms_coop_kernel "team kernel for big images (cfg5)"
hist yh_decode_nms_half.o 'decode_nms_tma_kernelILi2ELi20ELi2E6__half' decode_nms_tma_kernel_half "the tile kernel on a float16 head"
hist yh_loss.o 'loss_kernelILb1E' loss_kernel "loss forward+backward, TMA ring (small batches)"
hist yh_loss.o 'loss_gather_kernelILb1E' loss_gather_kernel "loss forward+backward, gather variant (cfg3 and larger)"
hist yh_map.o 'eval_update_kernelILb1E' eval_update_kernel "evaluator update: chained-scan append + per-image matching"
hist yh_map.o 'map_match_kernel' map_match_kernel "matching on arbitrary rows"
hist yh_map_reduce.o 'map_radix_kernel' map_radix_kernel "mAP reduce: persistent cooperative radix sort + AP"
hist yh_comm.o 'map_exchange_kernel' map_exchange_kernel "mAP exchange: peer stores + release flags"
hist yh_adapters.o 'encode_labels_kernel' encode_labels_kernel "label-grid encoder (TMA store)"
hist yh_loss.o 'loss_stream_kernelILb1E' loss_stream_kernel "loss forward+backward, stream variant: TMA in (UBLKCP.S.G), gradient tile composed in shared memory, TMA out (UBLKCP.G.S)"
hist yh_eval_fused.o 'eval_state_kernelILi2ELi20ELi2E' eval_state_kernel "update_state in one launch: decode + NMS of both tensors, matching, chained-scan append"
