#!/bin/bash
# make_golden.sh -- regenerates tests/golden/ from the UNMODIFIED reference backend
# (oracle/_ref/libsimbackend_cuda.so, built by oracle/Makefile.ref from /root/reference).
# Needs a GPU: run it through gpurun, then copy gpurun_out/golden/* into tests/golden/.
#   /usr/local/graft/bin/gpurun -- 'bash tests/golden/make_golden.sh'
# For every case: a reference state (positions + carried dynamic pattern), the first assembled
# Newton system of the next step (PPF_DUMP_LINSYS=0, main/dump_linsys.hpp), and the reference's
# result after that step (positions, dynamic-CSR pattern, per-iteration logs).
set -e
H=./ppf-contact-solver_b200/bin/ppf_host
REF=oracle/_ref/libsimbackend_cuda.so
export PPF_HOST_QUIET=1
gen() { # name scene n steps save_every states...
  name=$1; scene=$2; n=$3; steps=$4; every=$5; shift 5
  O=gpurun_out/golden/$name
  rm -rf $O; mkdir -p $O
  $H --lib $REF --scene $scene --n $n --steps $steps --save-every $every --out $O/traj > $O/traj.log 2>&1
  cases=""
  for st in "$@"; do
    cp $O/traj/state_$st.bin $O/state_$st.bin
    mkdir -p $O/ref_$st
    PPF_DUMP_LINSYS=0 PPF_DUMP_DIR=$O/ref_$st $H --lib $REF --scene $scene --n $n --steps 1 --load $O/state_$st.bin --out $O/ref_$st > $O/ref_$st.log 2>&1
    python - "$O/ref_$st" <<'PY'
import json, os, sys
d = sys.argv[1]
logs = {}
for ch in ("num_contact", "iter", "toi", "SL_toi", "SL_rod_toi", "max_sigma", "newton_steps", "toi_advanced", "final_dt", "max_dx", "reresid", "dyn_consumed"):
    p = os.path.join(d, "data", "advance.%s.out" % ch)
    logs[ch] = [float(l.split()[1]) for l in open(p)] if os.path.exists(p) else []
json.dump(logs, open(os.path.join(d, "logs.json"), "w"))
PY
    rm -rf $O/ref_$st/data $O/ref_$st/final_prev.bin $O/ref_$st/state_final.bin $O/ref_$st/steps.jsonl
    cases="$cases{\"state\": \"state_$st.bin\", \"dir\": \"ref_$st\"},"
  done
  rm -rf $O/traj $O/*.log
  echo "{\"scene\": \"$scene\", \"n\": $n, \"cases\": [${cases%,}]}" > $O/meta.json
}
case "$1" in
  five-only) gen five_twist_n24 five-twist 24 100 50 00050 00100 ;;
  hang-only) gen hang_n32 hang 32 300 150 00150 00300 ;;
  woven-only) gen woven_n32 woven 32 120 60 00060 00120 ;;
  woven-sl-only) gen woven_sl_n32 woven-sl 32 160 80 00080 00160 ;;
  drape-only) gen drape_n32 drape 32 140 70 00070 00140 ;;
  next-round) # cases that have no reference golden yet (features added after the GPU budget of round 1 was spent)
     ( gen headless_ref_n24 headless 24 68 68 00068 ) || echo "golden FAILED: headless_ref_n24"
     ( gen lift_n24 lift 24 40 40 00040 ) || echo "golden FAILED: lift_n24"
     ( gen sheet_arap_n24 sheet-arap 24 20 20 00020 ) || echo "golden FAILED: sheet_arap_n24"
     ( gen sheet_stvk_n24 sheet-stvk 24 20 20 00020 ) || echo "golden FAILED: sheet_stvk_n24"
     ( gen sheet_snhk_n24 sheet-snhk 24 20 20 00020 ) || echo "golden FAILED: sheet_snhk_n24"
     ( gen sheet_damped_n24 sheet-damped 24 20 20 00020 ) || echo "golden FAILED: sheet_damped_n24"
     ( gen woven_damped_n32 woven-damped 32 60 60 00060 ) || echo "golden FAILED: woven_damped_n32"
     ( gen seam_n16 seam 16 40 40 00040 ) || echo "golden FAILED: seam_n16"
     ( gen sheet_plastic_n24 sheet-plastic 24 20 20 00020 ) || echo "golden FAILED: sheet_plastic_n24"
     ( gen sheet_fixxz_n24 sheet-fixxz 24 20 20 00020 ) || echo "golden FAILED: sheet_fixxz_n24"
     ( gen drape_mesh_n24 drape-mesh 24 40 20 00020 00040 ) || echo "golden FAILED: drape_mesh_n24"
     ( gen sheet_torque_n24 sheet-torque 24 30 15 00015 00030 ) || echo "golden FAILED: sheet_torque_n24"
     ( gen balloon_n24 balloon 24 20 10 00010 00020 ) || echo "golden FAILED: balloon_n24"
     ( gen headless_cw_closed_n24 headless-cw-closed 24 68 68 00068 ) || echo "golden FAILED: headless_cw_closed_n24"
     ( gen fast_twist_quad_n16 fast-twist@quad 16 60 30 00030 00060 ) || echo "golden FAILED: fast_twist_quad_n16"
     ( gen fast_twist_log_n16 fast-twist@log 16 60 30 00030 00060 ) || echo "golden FAILED: fast_twist_log_n16" ;;
  *) gen fast_twist_n16 fast-twist 16 90 30 00030 00060 00090
     gen fast_twist_n24 fast-twist 24 80 40 00040 00080
     gen five_twist_n24 five-twist 24 100 50 00050 00100
     gen hang_n32 hang 32 300 150 00150 00300
     gen woven_n32 woven 32 120 60 00060 00120
     gen woven_sl_n32 woven-sl 32 160 80 00080 00160
     gen drape_n32 drape 32 140 70 00070 00140 ;;
esac
du -sh gpurun_out/golden/*
