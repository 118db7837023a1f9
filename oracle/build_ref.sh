#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.  Compiles pieces of the *unmodified* reference
# (aquaskyline/SOAP3-dp, mounted read-only at $REF, default /root/reference)
# from the sources where they lie into oracle/_ref/ (git-ignored, travels to
# the GPU box).  Nothing here is shipped or measured as the product.
#
# Outputs (all under oracle/_ref/):
#   soap3-dp-builder, BGS-Build (+ .ini)   reference index builders (Makefile:98-102)
#   libref_search.so   DV-Kernel.cu device code compiled for the host through
#                      oracle/ref_shim/cuda_host_shim.h (bit-exact answer slots)
#   libref_dp.so       DV-DPfunctions.cu:35-512 (DP kernels) compiled for the host
#   libref_dp_cuda.so  the same kernels compiled for sm_100a, launched as performAlignment launches them
#   libref_seed_pair.so  findRevStart + pairEndMerge of paired-end DP seeding (DV-DPfunctions.cu:2626-2653,2780-2880)
#   libref_decode.so   CigarStringEncoder + the result loop of algnmtCPUThread + convertToCigarStr (DV-DPfunctions.h:514-597, .cu:1699-1733, PE.cpp:83-110,420-483)
#   libref_pair.so     PEMappingOccurrences + PEStatsPEPairList and what they call (PEAlgnmt.cpp:114-361,480-637,777-838)
#   libref_windows.so  the DP engines' batch packers: SingleEndAlgnBatch::pack, HalfEndAlgnBatch::pack, packLeft / packRight (DV-DPfunctions.cu:1425,2027,3374,3420)
#   libref_retain.so   retainAllBest / retainAllBestWithCap / retainAllBestAndSecBest + list helpers (SAList.cpp:26-69,140-390)
#   libref_md.so       getMisInfoForDP (PE.cpp:499-666): MD string, mismatch / gap counts of a DP alignment
#   shim_check (+ shim_case/)  the drop-in shim linked with a driver written against the reference's headers, and its test case
#   libref_mapq.so     the nine MAPQ functions of the SAM writers + tables (BGS-IO.cpp:33-45,2280-2580)
#   libref_params.so   getSeedPositions (definitions.h:323-442) + getParameterFor*DP (CPUfunctions.cpp:46-260)
#   libref_cpu_search.so  the reference's own CPU search: ProcessReadDoubleStrand2 (CPUfunctions.cpp:555-622) + BGS-HostAlgnmtAlgo2.cpp,
#                      SAList.cpp, SRA2BWTMdl.c, SRA2BWTCheckAndExtend.c, BWT.c compiled from where they lie; the CPU arm of bench.py
#   libref_validate.so validateAlignments (CPUfunctions.cpp:1129-1222) + the packers / popcount distance it calls (PE.cpp:28-60,148-206,287-325)
#   libref_sam.so      pairOutputSAMAPI (BGS-IO.cpp:3478-3793) and all it calls: BGS-IO.cpp, PE.cpp, SAM.cpp, PEAlgnmt.cpp, SAList.cpp, samtools' bam_aux.c,
#                      whole; samwrite is the shim's and keeps the bam1_t records
#   libref_seed.so     the seed-hit radix sorts + singleMerge of single-end DP seeding (DV-DPfunctions.h:60-95, .cu:1101-1141)
#
# Three build fixes are applied to *copies* in oracle/_ref/patched/
# (SURVEY.md §8c): 2bwt-lib/BWT.c:424 pointer comparison,
# 2bwt-flex/LTConstruct.c BuildLookupTable falling off the end without return, and the same
# in three helpers of SAM.cpp (libref_sam.so only).
set -euo pipefail
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
if [ ! -d "$REF" ]; then
  echo "[build_ref] $REF not present; keeping prebuilt files in $OUT" >&2
  exit 0
fi
mkdir -p "$OUT/obj" "$OUT/patched"
CXX=${S3_CXX:-/usr/bin/g++}   # NB: the image exports CXX=/opt/gcc/bin/g++, which has no libgomp.spec
CFLAGS="-O3 -funroll-loops -fomit-frame-pointer -fpermissive -w -mpopcnt -fPIC"
JOBS=${JOBS:-8}

# ---- 2bwt-lib objects (Makefile BWTOBJLIBS) --------------------------------
sed '424s/bwt->cachedSaIndex > 0/bwt->cachedSaIndex != NULL/' "$REF/2bwt-lib/BWT.c" > "$OUT/patched/BWT.c"
sed 's/^\(\s*\)free(otop);/\1free(otop);\n\1return 0;/' "$REF/2bwt-flex/LTConstruct.c" > "$OUT/patched/LTConstruct.c"

BWTSRC="dictionary DNACount HSP HSPstatistic iniparser inistrlib karlin MemManager MiscUtilities QSufSort r250 TextConverter Timing Socket BWTConstruct"
CPUSRC="HOCC LT HOCCConstruct SRA2BWTCheckAndExtend SRA2BWTMdl"
pids=()
cc() { # src obj incdir
  if [ ! -f "$2" ] || [ "$1" -nt "$2" ]; then
    $CXX $CFLAGS -I"$3" -c "$1" -o "$2" &
    pids+=($!)
    if [ ${#pids[@]} -ge $JOBS ]; then wait "${pids[0]}"; pids=("${pids[@]:1}"); fi
  fi
}
for s in $BWTSRC; do cc "$REF/2bwt-lib/$s.c" "$OUT/obj/$s.o" "$REF/2bwt-lib"; done
cc "$OUT/patched/BWT.c" "$OUT/obj/BWT.o" "$REF/2bwt-lib"
for s in $CPUSRC; do cc "$REF/2bwt-flex/$s.c" "$OUT/obj/flex_$s.o" "$REF/2bwt-flex"; done
cc "$OUT/patched/LTConstruct.c" "$OUT/obj/flex_LTConstruct.o" "$REF/2bwt-flex"
wait
BWTOBJ=""; for s in dictionary DNACount HSP HSPstatistic iniparser inistrlib karlin MemManager MiscUtilities QSufSort r250 TextConverter Timing Socket BWT; do BWTOBJ="$BWTOBJ $OUT/obj/$s.o"; done
CPUOBJ=""; for s in $CPUSRC LTConstruct; do CPUOBJ="$CPUOBJ $OUT/obj/flex_$s.o"; done

if [ ! -x "$OUT/soap3-dp-builder" ]; then
  $CXX $CFLAGS -I"$REF/2bwt-flex" "$REF/2bwt-flex/2BWT-Builder.c" "$OUT/obj/BWTConstruct.o" $BWTOBJ $CPUOBJ -lm -o "$OUT/soap3-dp-builder"
  cp "$REF/soap3-dp-builder.ini" "$OUT/soap3-dp-builder.ini"
fi
if [ ! -x "$OUT/BGS-Build" ]; then
  $CXX $CFLAGS -I"$REF" "$REF/BGS-Build.cpp" $BWTOBJ -lm -o "$OUT/BGS-Build"
fi
echo "[build_ref] builders OK"

# ---- reference search kernels compiled for the host -------------------------
# A generated copy adds a rank-query counter (the roofline unit of SURVEY.md 8d)
# at the top of GPUBWTOccValue / GPUBWTAllOccValue / GPUBWTOccValueWithCumu.
# Second edit: `rightWord >> numLeftBits` with numLeftBits == 32 (read lengths whose middle
# base is the last of a word, e.g. 191) relies on the GPU's shift semantics (PTX shr/shl by
# >= 32 gives 0); on x86 the same C expression is undefined and yields rightWord.  The host
# copy spells the GPU result out so that it computes what the device computes.
sed -e 's/^\(\s*\)index -= ( index > inverseSa0 );/\1++s3_rank_queries; index -= ( index > inverseSa0 );/' \
    -e 's/( ( rightWord >> numLeftBits ) << numLeftBits )/( numLeftBits >= 32 ? 0u : ( ( rightWord >> numLeftBits ) << numLeftBits ) )/' \
    "$REF/DV-Kernel.cu" > "$OUT/patched/DV-Kernel.counted.cu"
$CXX -O2 -fopenmp -fpermissive -w -fPIC -shared -DS3_COUNT_RANK_QUERIES -I"$OUT/patched" -I"$REF" -I"$HERE/ref_shim" \
    "$HERE/ref_shim/ref_search_host.cpp" -o "$OUT/libref_search.so"
echo "[build_ref] libref_search.so OK"

# ---- reference DP kernels compiled for the host ------------------------------
# lines 35-512 of DV-DPfunctions.cu = _MAX/_MIN/_LOW_THRESHOLD, macros, DPScoreNHitPos,
# GenerateDPTable, SemiGlobalAligntment, GPUBacktrack.  The two texture references
# (removed from CUDA 12) become plain array reads; nothing else is touched.
sed -n '35,512p' "$REF/DV-DPfunctions.cu" \
  | sed -e '/^texture <uint>/d' \
        -e 's/tex1Dfetch(texPatterns, readTPARA + (((i)>>4)<<5))/X[readTPARA + (((i)>>4)<<5)]/' \
  > "$OUT/patched/dp_kernels.inc"
$CXX -O2 -fopenmp -fpermissive -w -fPIC -shared -I"$OUT/patched" -I"$HERE/ref_shim" \
    "$HERE/ref_shim/ref_dp_host.cpp" -o "$OUT/libref_dp.so"
echo "[build_ref] libref_dp.so OK"

# ---- reference seed-hit sort + merge (single-end DP seeding) on the host ------------------------
# radix sort macros, struct SeedPos, DPS_DIVIDE_GAP and the body of singleMerge, cut from where they lie; the method
# becomes a free function, nothing else is edited
{ sed -n '60,95p' "$REF/DV-DPfunctions.h"; sed -n '919,926p' "$REF/DV-DPfunctions.h"; sed -n '944p' "$REF/DV-DPfunctions.h" | sed 's/^\s*//';
  sed -n '1101,1141p' "$REF/DV-DPfunctions.cu" | sed 's/SingleEndSeedingEngine::SingleEndSeedingBatch::singleMerge/ref_singleMerge/'; } \
  > "$OUT/patched/seed_merge.inc"
$CXX -O2 -fpermissive -w -fPIC -shared -I"$OUT/patched" "$HERE/ref_shim/ref_seed_host.cpp" -o "$OUT/libref_seed.so"
echo "[build_ref] libref_seed.so OK"

# ---- reference paired-end seed-hit join on the host -----------------------------------------------
{ sed -n '60,95p' "$REF/DV-DPfunctions.h"; sed -n '1382,1393p' "$REF/DV-DPfunctions.h"; sed -n '1413p' "$REF/DV-DPfunctions.h" | sed 's/^\s*//';
  sed -n '2549p' "$REF/DV-DPfunctions.cu";
  sed -n '2626,2653p' "$REF/DV-DPfunctions.cu" | sed 's/PairEndSeedingEngine::PairEndSeedingBatch::findRevStart/ref_findRevStart/';
  sed -n '2780,2880p' "$REF/DV-DPfunctions.cu" | sed 's/PairEndSeedingEngine::PairEndSeedingBatch::pairEndMerge/ref_pairEndMerge/'; } \
  > "$OUT/patched/seed_pair.inc"
$CXX -O2 -fpermissive -w -fPIC -shared -I"$OUT/patched" "$HERE/ref_shim/ref_seed_pair_host.cpp" -o "$OUT/libref_seed_pair.so"
echo "[build_ref] libref_seed_pair.so OK"

# ---- reference DP result decoding (pattern bytes -> CIGAR, edit distance) on the host ---------------
{ sed -n '514,529p' "$REF/DV-DPfunctions.h"; sed -n '545,597p' "$REF/DV-DPfunctions.h"; } > "$OUT/patched/decode.inc"
sed -n '1699,1733p' "$REF/DV-DPfunctions.cu" > "$OUT/patched/decode_loop.inc"
{ sed -n '83,110p' "$REF/PE.cpp"; sed -n '420,485p' "$REF/PE.cpp" | sed 's/^int convertToCigarStr ( char \* special_cigar, char \* cigar, int \* deletedEnd )/int convertToCigarStr ( char * special_cigar, char * cigar, int * deletedEnd = NULL )/'; } > "$OUT/patched/sam_cigar.inc"
$CXX -O2 -fpermissive -w -fPIC -shared -I"$OUT/patched" "$HERE/ref_shim/ref_decode_host.cpp" -o "$OUT/libref_decode.so"
echo "[build_ref] libref_decode.so OK"

# ---- reference paired-end pairing (sort, merge walk, predicates, stats) against the reference's own header --------
{ sed -n '45,57p' "$REF/PEAlgnmt.cpp"; sed -n '114,361p' "$REF/PEAlgnmt.cpp"; sed -n '480,637p' "$REF/PEAlgnmt.cpp";
  sed -n '645,711p' "$REF/PEAlgnmt.cpp"; sed -n '777,838p' "$REF/PEAlgnmt.cpp"; } > "$OUT/patched/pair.inc"
$CXX -O1 -fpermissive -w -fPIC -shared -I"$OUT/patched" -I"$REF" -I"$REF/2bwt-lib" -I"$REF/2bwt-flex" \
    "$HERE/ref_shim/ref_pair_host.cpp" -o "$OUT/libref_pair.so"
echo "[build_ref] libref_pair.so OK"

# ---- reference MD-string builder of DP alignments against the reference's own header ---------------------------
{ sed -n '83,110p' "$REF/PE.cpp"; sed -n '499,666p' "$REF/PE.cpp"; } > "$OUT/patched/md.inc"
$CXX -O1 -fpermissive -w -fPIC -shared -I"$OUT/patched" -I"$REF" -I"$REF/2bwt-lib" -I"$REF/2bwt-flex" \
    "$HERE/ref_shim/ref_md_host.cpp" -o "$OUT/libref_md.so"
echo "[build_ref] libref_md.so OK"

# ---- reference MAPQ functions and their tables ------------------------------------------------------------
{ sed -n '33p;42p;45p' "$REF/BGS-IO.cpp"; sed -n '3014,3019p' "$REF/CPUfunctions.cpp"; sed -n '2280,2580p' "$REF/BGS-IO.cpp"; } > "$OUT/patched/mapq.inc"
$CXX -O2 -fpermissive -w -fPIC -shared -I"$OUT/patched" "$HERE/ref_shim/ref_mapq_host.cpp" -o "$OUT/libref_mapq.so"
echo "[build_ref] libref_mapq.so OK"

# ---- reference best-hit filters (retainAllBest family) against the reference's own header ---------------------
{ sed -n '26,69p' "$REF/SAList.cpp"; sed -n '140,390p' "$REF/SAList.cpp"; } > "$OUT/patched/retain.inc"
$CXX -O1 -fpermissive -w -fPIC -shared -I"$OUT/patched" -I"$REF" -I"$REF/2bwt-lib" -I"$REF/2bwt-flex" \
    "$HERE/ref_shim/ref_retain_host.cpp" -o "$OUT/libref_retain.so"
echo "[build_ref] libref_retain.so OK"

# ---- reference DP batch packers: which windows, clips, anchors the three engines hand to the kernel --------------
sed -n '1425,1468p' "$REF/DV-DPfunctions.cu" | sed 's/int SingleEndAlignmentEngine::SingleEndAlgnBatch::pack (/int ref_single_pack (/' > "$OUT/patched/win_single.inc"
sed -n '2027,2110p' "$REF/DV-DPfunctions.cu" | sed 's/int HalfEndAlignmentEngine::HalfEndAlgnBatch::pack (/int ref_half_pack (/' > "$OUT/patched/win_half.inc"
sed -n '3374,3472p' "$REF/DV-DPfunctions.cu" | sed 's/int PairEndAlignmentEngine::PairEndAlgnBatch::packLeft (/int ref_pair_packLeft (/; s/void PairEndAlignmentEngine::PairEndAlgnBatch::packRight ()/void ref_pair_packRight ()/' > "$OUT/patched/win_pair.inc"
$CXX -O1 -fpermissive -w -fPIC -shared -I"$OUT/patched" "$HERE/ref_shim/ref_windows_host.cpp" -o "$OUT/libref_windows.so"
echo "[build_ref] libref_windows.so OK"

# ---- reference stage tables (seed layout, per-stage DP parameters) against the reference's own headers --------
sed -n '46,260p' "$REF/CPUfunctions.cpp" > "$OUT/patched/params.inc"
$CXX -O1 -fpermissive -w -fPIC -shared -I"$OUT/patched" -I"$REF" -I"$REF/2bwt-lib" -I"$REF/2bwt-flex" \
    "$HERE/ref_shim/ref_params_host.cpp" -o "$OUT/libref_params.so"
echo "[build_ref] libref_params.so OK"

# ---- reference long-read validation of seed alignments, against the reference's own headers -------------------------------
{ sed -n '28,60p' "$REF/PE.cpp"; sed -n '148,206p' "$REF/PE.cpp"; sed -n '287,325p' "$REF/PE.cpp"; sed -n '1129,1222p' "$REF/CPUfunctions.cpp"; } > "$OUT/patched/validate.inc"
$CXX -O1 -fpermissive -w -fPIC -shared -I"$OUT/patched" -I"$REF" -I"$REF/2bwt-lib" -I"$REF/2bwt-flex" \
    "$HERE/ref_shim/ref_validate_host.cpp" -o "$OUT/libref_validate.so"
echo "[build_ref] libref_validate.so OK"

# ---- the reference's SAM writer of a properly paired read pair, records kept instead of written ------------------------------------
# SAM.cpp:58,70,80: three int functions fall off their end; g++ then drops the loop exits (undefined behaviour).  `return 0;` is added
# to a COPY (the third build fix of this kind, see the top of the script); everything else is compiled from where it lies.
sed -e '58s/^}/    return 0;\n}/' -e '70s/^}/    return 0;\n}/' -e '80s/^}/    return 0;\n}/' "$REF/SAM.cpp" > "$OUT/patched/SAM.cpp"
sed -n '24,41p' "$REF/samtools-0.1.18/bam_import.c" > "$OUT/patched/sam_nt16.inc"
sed -n '3014,3019p' "$REF/CPUfunctions.cpp" > "$OUT/patched/sam_bwase.inc"
# the text formatter of a record (what samwrite prints into a SAM file): bam_format1_core cut by line range, with samtools' own kstring
sed -n '243,324p' "$REF/samtools-0.1.18/bam.c" > "$OUT/patched/sam_format.inc"
/usr/bin/gcc -O1 -w -fPIC -I"$REF/samtools-0.1.18" -c "$REF/samtools-0.1.18/kstring.c" -o "$OUT/obj/sam_kstring.o"
for f in BGS-IO PE PEAlgnmt; do $CXX -O1 -fpermissive -w -fPIC -I"$REF" -I"$REF/2bwt-lib" -I"$REF/2bwt-flex" -c "$REF/$f.cpp" -o "$OUT/obj/sam_$f.o"; done
$CXX -O1 -fpermissive -w -fPIC -I"$REF" -I"$REF/2bwt-lib" -I"$REF/2bwt-flex" -c "$OUT/patched/SAM.cpp" -o "$OUT/obj/sam_SAM.o"
/usr/bin/gcc -O1 -w -fPIC -I"$REF/samtools-0.1.18" -c "$REF/samtools-0.1.18/bam_aux.c" -o "$OUT/obj/sam_bam_aux.o"
$CXX $CFLAGS -I"$REF" -I"$REF/2bwt-lib" -I"$REF/2bwt-flex" -c "$REF/SAList.cpp" -o "$OUT/obj/SAList.o"
$CXX -O1 -fpermissive -w -fPIC -shared -I"$OUT/patched" -I"$REF" -I"$REF/2bwt-lib" -I"$REF/2bwt-flex" -I"$REF/samtools-0.1.18" "$HERE/ref_shim/ref_sam_host.cpp" \
    "$OUT/obj/sam_BGS-IO.o" "$OUT/obj/sam_PE.o" "$OUT/obj/sam_SAM.o" "$OUT/obj/sam_PEAlgnmt.o" "$OUT/obj/SAList.o" "$OUT/obj/sam_bam_aux.o" "$OUT/obj/sam_kstring.o" $BWTOBJ -lm -o "$OUT/libref_sam.so"
echo "[build_ref] libref_sam.so OK"

# ---- the reference's CPU search path: models, lookup-table + BWT backward / bidirectional search, check-and-extend ------------
# ProcessReadDoubleStrand2 is cut by line range (CPUfunctions.cpp as a whole needs the aligner around it); the files it calls
# into are compiled whole and unmodified.  ref_cpu_search_host.cpp builds BWT / LT / HSP structs from arrays and loops over reads.
sed -n '555,622p' "$REF/CPUfunctions.cpp" > "$OUT/patched/cpu_search.inc"
$CXX $CFLAGS -I"$REF" -I"$REF/2bwt-lib" -I"$REF/2bwt-flex" -c "$REF/BGS-HostAlgnmtAlgo2.cpp" -o "$OUT/obj/HostAlgnmtAlgo2.o"
$CXX $CFLAGS -I"$REF" -I"$REF/2bwt-lib" -I"$REF/2bwt-flex" -c "$REF/SAList.cpp" -o "$OUT/obj/SAList.o"
$CXX $CFLAGS -fopenmp -shared -I"$OUT/patched" -I"$REF" -I"$REF/2bwt-lib" -I"$REF/2bwt-flex" "$HERE/ref_shim/ref_cpu_search_host.cpp" \
    "$OUT/obj/HostAlgnmtAlgo2.o" "$OUT/obj/SAList.o" "$OUT/obj/BWTConstruct.o" $BWTOBJ $CPUOBJ -lm -o "$OUT/libref_cpu_search.so"
echo "[build_ref] libref_cpu_search.so OK"

# ---- the same DP kernels compiled for sm_100a: the reference's GPU kernels on the B200 ("kernel to beat") -----
NVCC=${S3_NVCC:-/usr/local/cuda/bin/nvcc}
if [ -x "$NVCC" ]; then
  $NVCC -O3 -w -shared -Xcompiler -fPIC -ccbin "$CXX" -gencode arch=compute_100a,code=sm_100a -I"$OUT/patched" \
      "$HERE/ref_shim/ref_dp_cuda.cu" -o "$OUT/libref_dp_cuda.so" -lcudart
  echo "[build_ref] libref_dp_cuda.so OK"
fi

# ---- the reference-side shim compiles against the reference's unmodified headers, and runs --------------
# (integration/soap3dp_b200_shim.cpp = the file a SOAP3-dp maintainer links instead of the device code of alignment.cu /
# DV-DPfunctions.cu.)  shim_check is a small driver that calls it the way SOAP3-dp does -- Soap3Index / BWT / DPParameters
# from the reference's headers, GPUINDEXUpload, perform_round1_alignment, SemiGlobalAligner -- on the case
# oracle/make_shim_case.py writes, and compares with the oracle: oracle/_ref/shim_check oracle/_ref/shim_case (on a GPU box).
NVCC=${S3_NVCC:-/usr/local/cuda/bin/nvcc}
if [ -x "$NVCC" ]; then
  $NVCC -x cu -c -w -Xcompiler -fpermissive -ccbin "$CXX" -gencode arch=compute_100a,code=sm_100a \
      -I"$REF" -I"$REF/2bwt-lib" -I"$REF/2bwt-flex" -I"$HERE/../include" \
      "$HERE/../integration/soap3dp_b200_shim.cpp" -o "$OUT/obj/soap3dp_b200_shim.o"
  for sym in _Z14GPUINDEXUploadP10Soap3IndexPPjS2_S2_S2_ _Z24perform_round1_alignmentPjS_PA10_S_jjjjjbijyP10Soap3IndexS_S_S_S_ \
             _ZN17SemiGlobalAligner16performAlignmentEPjS0_S0_S0_PiS1_S0_S0_PhiS0_S0_S0_S0_ \
             _Z12alignSingleRPjS_S_jjP10Soap3IndexP16SingleAlignParamRyRjP16AlgnResultArrays \
             deepDPAlignResults singleDPAlignResults rescueDPAlignResults; do
    nm "$OUT/obj/soap3dp_b200_shim.o" > "$OUT/obj/soap3dp_b200_shim.nm"          # (not piped: grep -q would end nm early under pipefail)
    grep -q " T .*$sym" "$OUT/obj/soap3dp_b200_shim.nm" || { echo "[build_ref] shim lacks $sym" >&2; exit 1; }
  done
  echo "[build_ref] integration shim compiles against the reference headers"
  LIBDIR="$HERE/../soap3-dp_b200"
  if [ -f "$LIBDIR/libsoap3dp_b200.so" ]; then
    $NVCC -x cu -c -w -Xcompiler -fpermissive -ccbin "$CXX" -gencode arch=compute_100a,code=sm_100a \
        -I"$REF" -I"$REF/2bwt-lib" -I"$REF/2bwt-flex" -I"$HERE/../include" "$HERE/ref_shim/shim_check.cu" -o "$OUT/obj/shim_check.o"
    $CXX $CFLAGS -I"$REF" -I"$REF/2bwt-lib" -I"$REF/2bwt-flex" -c "$REF/global_arrays.cpp" -o "$OUT/obj/global_arrays.o"      # addOCCToArray, resultArraysConstruct
    $NVCC -Wno-deprecated-gpu-targets -ccbin "$CXX" -o "$OUT/shim_check" "$OUT/obj/shim_check.o" "$OUT/obj/soap3dp_b200_shim.o" "$OUT/obj/global_arrays.o" \
        -L"$LIBDIR" -lsoap3dp_b200 -Xlinker -rpath -Xlinker '$ORIGIN/../../soap3-dp_b200' -lcudart
    if [ ! -f "$OUT/shim_case/meta.txt" ] || [ ! -f "$OUT/shim_case/single_off.bin" ] || [ ! -f "$OUT/shim_case/deep_meta.txt" ] || [ ! -f "$OUT/shim_case/sdp_meta.txt" ] || [ ! -f "$OUT/shim_case/rescue_meta.txt" ]; then python "$HERE/make_shim_case.py"; fi
    echo "[build_ref] shim_check OK"
  fi
fi
