#!/usr/bin/env bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE: build the reference's own Fortran (UNMODIFIED sources, read where they lie
# under /root/reference) into oracle/_ref/ref_driver, for pinning the C++ oracle to the reference's bits.
#
#   * needs a Fortran compiler (gfortran | flang | nvfortran | lfortran).  Neither this image nor the gpurun box has one
#     (profiles/r02_fortran_probe_container.log, profiles/r02_fortran_probe_gpubox.log): the script then prints why and
#     exits 3, and the pin is taken from oracle/f90run instead (the reference's source text executed by the
#     Fortran-subset interpreter of this repository, tests/golden/make_golden_ref.py).
#   * the reference's mesh reader needs a modified cgnslib 3.2.1 + HDF5 (src/modules/mod_cgns.f90:4), which is neither
#     vendored nor installed, so the scratch copy of src/setup/cell_input.f90 gets its CGNS block (:36-93) replaced by
#     the raw-mesh reader cfd-lite_b200/fortran/mod_rawmesh.f90; a hook line is appended after the
#     two `write(*,oformat)` statements of src/modules/mod_solver.f90 so that the residual history leaves with full
#     precision.  Nothing else is edited; no reference source enters the repository (the scratch copy lives in $TMP).
#   * outputs go to oracle/_ref/ only (git-ignored, travels to the GPU box with the built libraries).
#
# usage: oracle/build_ref.sh [reference-root]      then: tests/test_oracle_vs_ref.py picks the binary up
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${1:-/root/reference}"
OUT="$HERE/_ref"
FC=""
for c in gfortran gfortran-14 gfortran-13 gfortran-12 flang-new flang nvfortran lfortran ifx; do
  if command -v "$c" >/dev/null 2>&1; then FC="$c"; break; fi
done
if [ -z "$FC" ]; then
  echo "build_ref.sh: no Fortran compiler on PATH (tried gfortran, flang, nvfortran, lfortran, ifx): oracle/_ref not built" >&2
  exit 3
fi
[ -d "$REF/src" ] || { echo "build_ref.sh: $REF/src not found" >&2; exit 3; }
case "$FC" in
  gfortran*) FFLAGS="-O3 -cpp -fdefault-real-8 -fdefault-double-8 -ffree-line-length-512 -ffp-contract=off -fno-fast-math" ;;
  nvfortran) FFLAGS="-O3 -Mpreprocess -r8 -Mfree -Mnofma -Kieee" ;;
  ifx)       FFLAGS="-O3 -fpp -r8 -free -fp-model=strict -no-fma" ;;
  *)         FFLAGS="-O3 -cpp -fdefault-real-8" ;;
esac
TMP="$(mktemp -d)"; trap 'rm -rf "$TMP"' EXIT
mkdir -p "$OUT"
S="$REF/src"
# scratch copies of the two files that need the edits described above
python3 - "$S" "$TMP" <<'PY'
import re, sys
S, T = sys.argv[1], sys.argv[2]
src = open(S + "/setup/cell_input.f90").read().split("\n")
# lines 36-93 (1-based): `cg%file=cgfilename` ... `call cgns_cell_data(...)`  ->  rawmesh_read + the assignments of :72-83
a = next(i for i, l in enumerate(src) if "cg%file=cgfilename" in l)
b = next(i for i, l in enumerate(src) if "call cgns_cell_data(" in l)
new = ["       call rawmesh_read(cgfilename, mg_lvl%cellname, nvx, nsec, mg_lvl%esec, mg_lvl%etype, mg_lvl%ne2vx, &",
       "                         mg_lvl%sectionName, mg_lvl%e2vx, ne2vxmax, ne, nf, nbf, vx2e_size, geom%x, geom%y, geom%z)",
       "       nelem=ne+nbf", "       nbndry = nbf", "       mg_lvl%nsec=nsec", "       mg_lvl%ne2vx_max=ne2vxmax", "       mg_lvl%nvx=nvx",
       "       mg_lvl%nelem=nelem", "       mg_lvl%nbndry=nbndry", "       mg_lvl%nfaces=nf", "       npmax=5"]
out = src[:a] + new + src[b + 1:]
txt = "\n".join(out)
txt = txt.replace("use mod_cgns", "use mod_rawmesh").replace("type (cgnsdb) :: cg", "")
txt = re.sub(r"\n\s*call cg_close_f\(cg%unit, ier\)\s*\n\s*if \(ier \.eq\. ERROR\) call cg_error_exit_f", "\n", txt)
open(T + "/cell_input.f90", "w").write(txt)
sol = open(S + "/modules/mod_solver.f90").read()
sol = sol.replace("  use mod_subdomains\n", "  use mod_subdomains\n  use ref_hist\n", 1)
sol = sol.replace("    write(*,oformat) name,it,res_i_tot,res_f_tot,res_max_tot\n", "    write(*,oformat) name,it,res_i_tot,res_f_tot,res_max_tot\n    call ref_hist_push(it,res_i_tot,res_f_tot,res_max_tot)\n")
sol = sol.replace("    write(*,oformat) name,it,res_i,res_f,res_max\n", "    write(*,oformat) name,it,res_i,res_f,res_max\n    call ref_hist_push(it,res_i,res_f,res_max)\n")
open(T + "/mod_solver.f90", "w").write(sol)
PY
# module order: util -> mesh data structures -> agglomeration -> geometry -> rawmesh -> cell_input -> properties ->
# equation base -> subdomains -> (history hook) -> solver -> equations -> physics -> driver
awk '/^module ref_hist/,/^end module/' "$HERE/ref/ref_driver.f90" > "$TMP/ref_hist.f90"
awk '/^program ref_driver/,/^end program/' "$HERE/ref/ref_driver.f90" > "$TMP/ref_main.f90"
SRCS=("$S/modules/mod_util.f90" "$S/setup/mod_meshds_uns.f90" "$S/setup/mod_agglomeration.f90" "$S/setup/mod_mg_lvl_uns.f90"
      "$S/setup/calc_aip_xyzip.f90" "$S/setup/calc_vol_cv_centers.f90" "$HERE/../cfd-lite_b200/fortran/mod_rawmesh.f90" "$TMP/cell_input.f90"
      "$S/modules/mod_properties.f90" "$S/modules/mod_eqn_setup.f90" "$S/modules/mod_subdomains.f90" "$TMP/ref_hist.f90" "$TMP/mod_solver.f90"
      "$S/equations/mod_scalar.f90" "$S/equations/mod_energy.f90" "$S/modules/mod_multiphase.f90" "$S/equations/mod_uvwp.f90"
      "$S/modules/mod_physics.f90" "$TMP/ref_main.f90")
( cd "$TMP" && ulimit -s unlimited 2>/dev/null || true; $FC $FFLAGS -J"$TMP" -I"$TMP" "${SRCS[@]}" -o "$OUT/ref_driver" )
echo "$FC $FFLAGS" > "$OUT/BUILD_INFO"
# the ISO_C_BINDING bridge of the product is compile-checked against the same modules
$FC $FFLAGS -J"$TMP" -I"$TMP" -c "$HERE/../cfd-lite_b200/fortran/mod_gpu_bridge.f90" -o "$TMP/mod_gpu_bridge.o" && echo "mod_gpu_bridge.f90: compiles" >> "$OUT/BUILD_INFO" \
  || echo "mod_gpu_bridge.f90: DOES NOT COMPILE" >> "$OUT/BUILD_INFO"
echo "built $OUT/ref_driver with $FC"
