"""Opcode census of the production kernels in the shipped libpa_b200.so (cuobjdump -sass), for profiles/.
What it shows: TMA bulk copies (UBLKCP) and mbarriers (SYNCS) in k_spmv_tma; fp64 DMUL/DADD without DFMA in the SpMV (the
reference's unfused `bi += aij*xj`); system-scope loads/stores (peer-mapped flags, scalars, ghost values) in the kernels that
carry the exchange and the all-reduce; DFMA only where it belongs (||r||^2 accumulation, the IEEE division routine)."""
import collections
import re
import subprocess
import sys

SO = sys.argv[1] if len(sys.argv) > 1 else "partitionedarrays.jl_b200/lib/libpa_b200.so"
KERNELS = [
    ("k_spmv_pat<int32 rowptr, BATCH 8, 2 rows per thread, fixed length 7> (7-pt 512^3, row patterns)", "_Z10k_spmv_patIiLi8ELi2ELi2ELi7EEv8SpmvArgsIT_E6TmaCfg"),
    ("k_spmv_pat<int64 rowptr, BATCH 16, 1 row per thread, fixed length 27> (27-pt 512^3, row patterns)", "_Z10k_spmv_patIlLi16ELi1ELi2ELi27EEv8SpmvArgsIT_E6TmaCfg"),
    ("k_spmv_tma<int32 rowptr, MODE 0 (local), BATCH 8> (column stream)", "_Z10k_spmv_tmaIiLi0ELi8ELb0EEv8SpmvArgsIT_E6TmaCfg"),
    ("k_spmv_tma<int64 rowptr, MODE 0, BATCH 16> (27-pt 512^3, column stream)", "_Z10k_spmv_tmaIlLi0ELi16ELb0EEv8SpmvArgsIT_E6TmaCfg"),
    ("k_spmv_tma<int32, MODE 4 (consistent! fused in), BATCH 8>", "_Z10k_spmv_tmaIiLi4ELi8ELb0EEv8SpmvArgsIT_E6TmaCfg"),
    ("k_gs_color_tma<27, batch 27, row patterns> (multi-colour Gauss-Seidel through the TMA ring)", "_Z14k_gs_color_tmaILi27ELi27ELb1EEv10GsSellArgsii"),
    ("k_residual_restrict<int64> (residual at the injection points)", "_Z19k_residual_restrictIlEvPdPKdS2_PKT_PKiS2_lllll"),
    ("k_consistent_sync (signal + wait + gather + done)", "_Z17k_consistent_syncPdPKiS1_S1_l8PeerPtrsPy8FlagPtrsS4_S4_iPjPi"),
    ("k_cg_direction_xchg (u = r + beta*u with consistent!(u) inside)", "_Z19k_cg_direction_xchgPdPKdl7RedWaitS_PKi8XchgArgs"),
    ("k_cg_update_fold (x, r update + ||r||^2 + folded all-reduce)", "_Z16k_cg_update_foldPdPKdS_S1_ll7RedWait8DoneWait7RedPushS1_PiS_Pj"),
    ("k_gs_flow_pipe<int64, 8 lanes per row> (bit-exact wavefront Gauss-Seidel)", "_Z14k_gs_flow_pipeIlLi8EEv6GsArgsIT_E"),
    ("k_gs_sell<27, MODE 0> (multi-colour Gauss-Seidel, one launch per colour)", "_Z9k_gs_sellILi27ELi0EEv10GsSellArgs"),
]
CLASSES = [("TMA / mbarrier", r"^(UBLKCP|SYNCS|UTMA)"), ("fp64", r"^(DMUL|DADD|DFMA|DSETP|MUFU\.RCP64H)"), ("global memory", r"^(LDG|STG|ATOMG|REDG|RED|ATOM|LD\.|ST\.)"),
           ("shared memory", r"^(LDS|STS)"), ("fences / barriers", r"^(MEMBAR|FENCE|BAR|WARPSYNC|ERRBAR|CCTL)")]

print(f"# opcode census, {SO} (sm_100a)")
for title, sym in KERNELS:
    out = subprocess.run(["cuobjdump", "-sass", "-fun", sym, SO], capture_output=True, text=True).stdout
    ops = collections.Counter()
    for line in out.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            ops[m.group(1)] += 1
    print(f"\n## {title}\n#  {sym}: {sum(ops.values())} instructions")
    if not ops:
        print("   (not found)")
        continue
    for cname, rx in CLASSES:
        sel = sorted(((n, o) for o, n in ops.items() if re.match(rx, o)), reverse=True)
        print(f"   {cname:18s}: " + (", ".join(f"{o} x{n}" for n, o in sel) if sel else "-"))
