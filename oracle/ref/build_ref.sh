#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.  Builds oracle/_ref/libref.so: the REFERENCE'S OWN kernels for the non-SGM
# stages, compiled for sm_100a straight from the sources where they lie under /root/reference
# (line ranges per SURVEY.md Appendix E) behind a type shim, plus our host harness (*.inc).
# The assembled translation units live in a temporary directory and are deleted; only the shared
# library lands in oracle/_ref/ (git-ignored, ships to the GPU box).  The SGM stage cannot be built:
# it is the third-party cv::cuda::StereoSGM (OpenCV-CUDA, absent here).
set -euo pipefail
R=${REFERENCE_ROOT:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT="$HERE/../_ref"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
[ -d "$R" ] || { echo "reference tree $R not present: keeping the prebuilt $OUT/libref.so"; exit 0; }
mkdir -p "$OUT"
if [ -f "$OUT/libref.so" ] && [ -f "$OUT/libref_host.so" ] && [ "$OUT/libref.so" -nt "$HERE/build_ref.sh" ] && [ "$OUT/libref.so" -nt "$HERE/shim.h" ] \
   && [ "$OUT/libref_host.so" -nt "$HERE/build_ref.sh" ] && [ "$OUT/libref_host.so" -nt "$HERE/shim_host.h" ] \
   && [ -z "$(find "$HERE" -name '*.inc' -newer "$OUT/libref.so")" ] && [ -z "$(find "$HERE" -name '*.inc' -newer "$OUT/libref_host.so")" ]; then exit 0; fi
T=$(mktemp -d)
trap 'rm -rf "$T"' EXIT
cut_() { sed -n "$2,$3p" "$R/$1"; }
CU=include/utils/cuda.cuh
SP=src/modules/superpixels/contourrelaxation
SPI=include/modules/superpixels/contourrelaxation
inc() { echo "#include \"$HERE/shim.h\""; }

{ inc; cut_ $CU 10 15; echo "namespace cart {"; cut_ $CU 59 191; echo "}";
  cut_ src/modules/disparity/derivative.cu 11 116; cat "$HERE/harness_derivative.inc"; } > "$T/derivative.cu"

{ inc; cut_ $CU 10 15; cut_ $CU 31 36; echo "namespace cart {"; cut_ $CU 47 51; cut_ $CU 59 191;
  cut_ include/modules/planeseg.hpp 25 41; echo "}"; echo "#define CARTSLAM_PLANE_COUNT 3";
  cut_ src/modules/planeseg/planeseg.cu 11 243; cat "$HERE/harness_naive.inc"; } > "$T/naive.cu"

{ inc; cut_ $CU 10 15; cut_ $CU 31 36; echo "namespace cart {"; cut_ $CU 47 51;
  cut_ include/modules/planeseg.hpp 25 41; echo "}"; echo "#define CARTSLAM_PLANE_COUNT 3";
  cut_ src/modules/planeseg/sp_planeseg.cu 12 21; cut_ src/modules/planeseg/sp_planeseg.cu 25 184;
  cat "$HERE/harness_sp_planeseg.inc"; } > "$T/sp_planeseg.cu"

{ inc; cut_ $CU 10 15; echo "namespace cart {"; cut_ $CU 59 191; echo "}";
  cut_ src/modules/disparity/interpolation.cu 8 82; cat "$HERE/harness_interpolate.inc"; } > "$T/interpolate.cu"

{ inc; cut_ $CU 10 15; cut_ $CU 25 27; echo "namespace cart {"; cut_ $CU 59 191; echo "}";
  echo "namespace cart::contour { double const featuresMinVariance = 1.0 / 12.0;";
  cut_ $SPI/features/ifeature.hpp 10 13; cut_ $SPI/features/feature.cuh 11 45; echo "}";
  cut_ $SPI/features/gaussian.cuh 8 10; echo "namespace cart::contour {"; cut_ $SPI/features/gaussian.cuh 14 69;
  cut_ $SP/features/gaussian.cu 5 21; cut_ $SP/features/gaussian.cu 30 206;
  cut_ $SPI/features/compactness.cuh 28 58; cut_ $SP/features/compactness.cu 5 20; cut_ $SP/features/compactness.cu 28 197;
  echo "}"; cut_ $SP/contourrelaxation.cu 10 327; cat "$HERE/harness_contour.inc"; } > "$T/contour.cu"

{ inc; cut_ $CU 10 15; cut_ $CU 31 36; echo "namespace cart {"; cut_ include/modules/planefit.hpp 18 23; echo "}";
  cut_ src/modules/planefit.cu 14 138; cat "$HERE/harness_planefit.inc"; } > "$T/planefit.cu"

{ inc; cut_ $CU 10 15; echo "namespace cart {"; cut_ include/modules/planeseg.hpp 36 66; echo "}";
  cut_ src/modules/planeseg/planeseg_vis.cu 14 56; cat "$HERE/harness_overlay.inc"; } > "$T/overlay.cu"

{ inc; cut_ $CU 10 15; cut_ src/modules/superpixels/visualization.cu 4 42; cat "$HERE/harness_overlay_sp.inc"; } > "$T/overlay_sp.cu"

FLAGS="-gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo --expt-relaxed-constexpr -Xcompiler -fPIC -w"
for n in derivative naive sp_planeseg interpolate contour planefit overlay overlay_sp; do
  $NVCC $FLAGS -c "$T/$n.cu" -o "$T/$n.o"
done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT/libref.so" "$T"/*.o -lcudart
echo "built $OUT/libref.so"

# The reference's HOST code for the plane parameters (persistence peak finder + histogram-peak provider), compiled
# verbatim with g++ behind shim_host.h: runs on any CPU, pins the oracle's orc_find_peaks / orc_histogram_peak_update.
{ echo "#include \"$HERE/shim_host.h\""; cut_ include/utils/peaks.hpp 8 22; echo; cut_ src/utils/peaks.cpp 3 73; echo;
  echo "namespace cart {"; cut_ src/modules/planeseg/planeseg.cu 404 458; echo "}"; cat "$HERE/harness_host.inc"; } > "$T/host_params.cpp"
# ... and its KITTI calibration line parser (src/sources/kitti.cpp:11-12 camera ids, :18-85 addLeadingZeros,
# KITTICameraCalibration, readLine) for the on-disk source of the host layer
{ echo "#include <cstdint>"; echo "#include <string>"; echo "#include <algorithm>"; cut_ src/sources/kitti.cpp 11 12; cut_ src/sources/kitti.cpp 18 85;
  cat "$HERE/harness_kitti.inc"; } > "$T/host_kitti.cpp"
${CXX:-g++} -O2 -std=c++17 -fPIC -w -shared -o "$OUT/libref_host.so" "$T/host_params.cpp" "$T/host_kitti.cpp"
echo "built $OUT/libref_host.so"
