#!/usr/bin/env bash
# Builds the UNMODIFIED reference C path (hartallo) into oracle/_ref/ from the sources where they lie under
# $HL_REFERENCE (default /root/reference).  Nothing from the reference is copied into the repository: the two
# one-line, non-algorithmic patches (decoder bit-reader inline asm; rdtsc timer) are applied with sed to temporary
# copies that are deleted when the build ends.  Recipe = SURVEY.md Appendix A.
#
# TEST INFRASTRUCTURE ONLY: the product (hartallo_b200/) never links or executes anything produced here.
set -euo pipefail
REF="${HL_REFERENCE:-/root/reference}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
JOBS="${JOBS:-$(nproc)}"
if [ ! -d "$REF/source/h264" ]; then
  echo "build_ref: reference tree not found at $REF (GPU box uses the prebuilt oracle/_ref)" >&2
  exit 0
fi
mkdir -p "$OUT"
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
mkdir -p "$TMP/inc/hartallo/h264" "$TMP/shim" "$TMP/obj" "$TMP/src"
# patch 1: decoder-only bit reader uses malformed GNU inline asm -> force the portable branch
sed '203s/#if defined(__GNUC__)/#if 0/' "$REF/include/hartallo/h264/hl_codec_264_bits.h" > "$TMP/inc/hartallo/h264/hl_codec_264_bits.h"
# patch 2: timing helper only ('#error "Not implemented: use rdtsc inline asm"')
sed '376s/.*/    return __builtin_ia32_rdtsc();/' "$REF/source/hl_cpu.c" > "$TMP/src/hl_cpu.c"
# workaround 1: C files include <cfloat>
printf '#include <float.h>\n#include <limits.h>\n' > "$TMP/shim/cfloat"
# workaround 2: struct tags first declared inside prototypes + missing <limits.h>
{ printf '#include <limits.h>\n#include <float.h>\n'; \
  grep -rhoE 'struct +hl_[a-z0-9_]+' "$REF/include" "$REF/source" | sed -E 's/struct +/struct /' | sort -u | sed 's/$/;/'; } > "$TMP/fwd.h"
CF="-std=gnu99 -O2 -w -fPIC -fcommon -D_GNU_SOURCE -DHL_DISABLE_INTRIN=1 -DHL_DISABLE_ASM=1 -include $TMP/fwd.h -I$TMP/inc -I$REF/include -I$TMP/shim"
SRCS=$(ls "$REF"/source/*.c "$REF"/source/h264/*.c | grep -v /test | grep -v hl_codec_264_me.c | grep -v hl_x86_globals.c | grep -v '/hl_cpu.c$')
SRCS="$SRCS $TMP/src/hl_cpu.c"
echo "$SRCS" | tr ' ' '\n' | grep -v '^$' | xargs -P "$JOBS" -I{} sh -c 'gcc '"$CF"' -c "$1" -o '"$TMP"'/obj/$(basename "${1%.c}").o' _ {}
rm -f "$OUT/libhartallo_ref.a"
ar rcs "$OUT/libhartallo_ref.a" "$TMP"/obj/*.o
# shared object with every reference symbol visible: lets tests call the *_cpp kernels directly (ctypes)
gcc -shared -o "$OUT/libhartallo_ref.so" -Wl,--whole-archive "$OUT/libhartallo_ref.a" -Wl,--no-whole-archive -lpthread -lm -ldl
# driver (our code, reference public API + link-time wrappers that record per-MB decisions / per-candidate traces)
WRAPS="-Wl,--wrap=hl_codec_264_interpol_luma -Wl,--wrap=hl_codec_264_residual_write_block_cavlc -Wl,--wrap=hl_codec_264_rdo_mb_guess_best_inter_pred_avc -Wl,--wrap=hl_codec_264_rdo_mb_guess_best_intra_pred_avc -Wl,--wrap=hl_codec_264_me_ds_mb_find_best_cost -Wl,--wrap=hl_codec_264_nal_slice_data_encode -Wl,--wrap=hl_codec_264_rdo_mb_guess_best_inter_pred_svc -Wl,--wrap=hl_codec_264_rdo_mb_guess_best_intra_pred_svc"
if [ -f "$HERE/ref_driver.c" ]; then
  gcc $CF -c "$HERE/ref_driver.c" -o "$TMP/obj_driver.o"
  gcc "$TMP/obj_driver.o" $WRAPS "$OUT/libhartallo_ref.a" -lpthread -lm -ldl -o "$OUT/hl_ref_driver"
fi
# kernel-level harness: reference kernels behind plain-pointer entry points (ctypes)
if [ -f "$HERE/ref_kernels.c" ]; then
  gcc $CF -c "$HERE/ref_kernels.c" -o "$TMP/obj_kernels.o"
  gcc -shared -o "$OUT/libref_kernels.so" "$TMP/obj_kernels.o" -Wl,--whole-archive "$OUT/libhartallo_ref.a" -Wl,--no-whole-archive -lpthread -lm -ldl
fi
# the reference host code with the B200 hot path dropped in (host/hlb200_glue.c + libhl_b200.so): used by the bitstream-MD5 parity test
GLUE="$HERE/../host/hlb200_glue.c"
if [ -f "$GLUE" ] && [ -f "$HERE/ref_driver.c" ]; then
  gcc $CF -DHL_DRIVER_NO_WRAPS -c "$HERE/ref_driver.c" -o "$TMP/obj_driver_nw.o"
  gcc $CF -I"$HERE/../include" -c "$GLUE" -o "$TMP/obj_glue.o"
  GS="-Wl,--wrap=hl_codec_264_rdo_mb_guess_best_inter_pred_svc -Wl,--wrap=hl_codec_264_rdo_mb_guess_best_intra_pred_svc"
  GW="-Wl,--wrap=hl_codec_264_nal_slice_data_encode -Wl,--wrap=hl_codec_264_rdo_mb_guess_best_inter_pred_avc -Wl,--wrap=hl_codec_264_rdo_mb_guess_best_intra_pred_avc $GS"
  # libhl_b200.so is resolved at run time relative to the binary (oracle/_ref -> hartallo_b200)
  gcc "$TMP/obj_driver_nw.o" "$TMP/obj_glue.o" $GW "$OUT/libhartallo_ref.a" -L"$HERE/../hartallo_b200" -lhl_b200 -Wl,-rpath,'$ORIGIN/../../hartallo_b200' -lpthread -lm -ldl -o "$OUT/hl_b200_encoder" \
    || echo "build_ref: hl_b200_encoder not linked (build hartallo_b200/libhl_b200.so first)" >&2
  # the multi-stream drop-in: many codec instances through hl_codec_encode, ONE device launch per picture of all of them (host/hl_b200_multi.c, glue batch mode)
  if [ -f "$HERE/../host/hl_b200_multi.c" ]; then
    gcc $CF -I"$HERE/../include" -I"$HERE/../host" -c "$HERE/../host/hl_b200_multi.c" -o "$TMP/obj_multi.o"
    gcc "$TMP/obj_multi.o" "$TMP/obj_glue.o" $GW "$OUT/libhartallo_ref.a" -L"$HERE/../hartallo_b200" -lhl_b200 -Wl,-rpath,'$ORIGIN/../../hartallo_b200' -lpthread -lm -ldl -o "$OUT/hl_b200_multi" \
      || echo "build_ref: hl_b200_multi not linked" >&2
  fi
fi
# CPU check of the SVC enhancement-layer hook of the glue: base layer on the reference's CPU path, enhancement layers through the glue with the device source
# compiled as C++ standing in for libhl_b200.so (tools/emu/svc_shim.cpp + svc_emu.cpp)
if [ -f "$GLUE" ] && [ -f "$HERE/../tools/emu/svc_shim.cpp" ]; then
  gcc $CF -I"$HERE/../include" -DHLB200_GLUE_HOST_BASE_LAYER -c "$GLUE" -o "$TMP/obj_glue_svc.o"
  g++ -std=c++17 -O2 -w -fPIC -x c++ -c "$HERE/../tools/emu/svc_shim.cpp" -o "$TMP/obj_shim.o"
  g++ -std=c++17 -O2 -w -fPIC -x c++ -c "$HERE/../tools/emu/svc_emu.cpp" -o "$TMP/obj_emu.o"
  g++ "$TMP/obj_driver_nw.o" "$TMP/obj_glue_svc.o" "$TMP/obj_shim.o" "$TMP/obj_emu.o" -Wl,--wrap=hl_codec_264_nal_slice_data_encode $GS "$OUT/libhartallo_ref.a" -lpthread -lm -ldl -o "$OUT/hl_svc_glue_check"
  # the WHOLE glue (base layer hook included, exactly the object linked into hl_b200_encoder) with the per-macroblock sources of BOTH kernels compiled as C++
  # standing in for the library: every layer goes through the drop-in path, nothing through the reference's decision functions
  g++ -std=c++17 -O2 -w -fPIC -ffp-contract=off -DSVC_SHIM_WITH_SLICE -x c++ -c "$HERE/../tools/emu/svc_shim.cpp" -o "$TMP/obj_shim_full.o"
  g++ "$TMP/obj_driver_nw.o" "$TMP/obj_glue.o" "$TMP/obj_shim_full.o" "$TMP/obj_emu.o" $GW "$OUT/libhartallo_ref.a" -lpthread -lm -ldl -o "$OUT/hl_glue_check_full"
  # CPU check of the multi-stream driver + the glue's batch mode (same stand-in for the library)
  if [ -f "$TMP/obj_multi.o" ]; then
    g++ "$TMP/obj_multi.o" "$TMP/obj_glue.o" "$TMP/obj_shim_full.o" "$TMP/obj_emu.o" $GW "$OUT/libhartallo_ref.a" -lpthread -lm -ldl -o "$OUT/hl_multi_check"
  fi
fi
echo "build_ref: built $(ls "$OUT")"
