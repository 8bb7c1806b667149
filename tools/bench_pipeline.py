#!/usr/bin/env python3
"""Pipeline-level timing of BASELINE configs[0] (SURVEY 8d "as-shipped mode"): the 272 codewords of one decoder.py run.

  reference, as shipped : one `ldpc_ref` process per codeword, serially, each re-parsing the .pchk and two text files
                          (ex_decoder/decoder.py:553-558 -> def_func.py:47-51), on one host core like os.system does
  this repo             : ONE `ldpc ... --list frames.lst` process: inputs parsed in parallel while the
                          CUDA context comes up, one batched GPU call, 272 dec_*.txt written
  this repo, unmodified : the pipeline's own loop, one `ldpc` process per codeword, serially - each a thin client of a
  pipeline                resident worker (`ldpc --serve`, $DNALDPC_SOCKET) that keeps the CUDA context and the decoder

Inputs: the 272 true codewords through a BSC with flip probability --flip, LLR = +-ln((1-eps)/eps) with the pipeline's
eps = 0.02 (decoder.py --epsil), max_iter 200 (def_func.py:49). Every dec_*.txt of the two runs is compared byte for byte.
Prints one JSON object. Needs oracle/_ref/ldpc_ref (prebuilt) and a GPU.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oraclelib as ol  # noqa: E402

LDPC = os.path.join(ROOT, "dna-ldpc-codes_b200", "ldpc")
REF = os.path.join(ROOT, "oracle", "_ref", "ldpc_ref")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=272)
    ap.add_argument("--flip", type=float, default=0.006)
    ap.add_argument("--eps", type=float, default=0.02)
    ap.add_argument("--max-iter", type=int, default=200)
    ap.add_argument("--ref-frames", type=int, default=272, help="how many of the frames the serial reference loop runs (scaled up in the report)")
    a = ap.parse_args()
    cws = ol.load_codewords()
    N = 18432
    L = np.log((1 - a.eps) / a.eps)
    out = {"frames": a.frames, "flip": a.flip, "eps": a.eps, "max_iter": a.max_iter}
    with tempfile.TemporaryDirectory() as tmp:
        dirs = {}
        for who in ("ours", "ref", "worker"):
            d = os.path.join(tmp, who)
            os.makedirs(d)
            shutil.copyfile(ol.PCHK_18432, os.path.join(d, "decode_n18432_m2048_final.pchk"))
            dirs[who] = d
        names = []
        for f in range(a.frames):
            cw = cws[f % 272]
            recv = cw ^ ol.bsc_flips(7, f, N, a.flip)
            llr = np.where(recv == 0, L, -L)
            cwn, sn = "codeword_n18432_m1860_%d" % (f + 1), "soft72000_n18432_m1860_%d" % (f + 1)
            cw_txt = "".join("%d " % b for b in cw)
            soft_txt = " ".join(repr(float(x)) for x in llr)
            for d in dirs.values():
                open(os.path.join(d, cwn + ".txt"), "w").write(cw_txt)
                open(os.path.join(d, sn + ".txt"), "w").write(soft_txt)
            names.append((cwn, sn))
        with open(os.path.join(dirs["ours"], "frames.lst"), "w") as fh:
            fh.write("".join("%s %s\n" % n for n in names))
        # ours: one process
        t0 = time.perf_counter()
        r = subprocess.run([LDPC, "0", "0", "0", "7", str(a.max_iter), "1", "x", "x", "decode_n18432_m2048_final", "0", "0", "0", "0",
                            "--list", "frames.lst", "--timing"], cwd=dirs["ours"], capture_output=True, text=True)
        out["ours_one_process_s"] = time.perf_counter() - t0
        if r.returncode != 0:
            raise SystemExit("ldpc --list failed: " + r.stderr)
        out["ours_timing"] = json.loads(r.stderr.strip().splitlines()[-1])
        # ours, process per codeword like the unmodified pipeline, against a resident worker
        sock = os.path.join(tmp, "ldpc.sock")
        env = dict(os.environ, DNALDPC_SOCKET=sock)
        worker = subprocess.Popen([LDPC, "--serve"], cwd="/", env=env, stderr=subprocess.DEVNULL)
        try:
            for _ in range(200):
                if os.path.exists(sock):
                    break
                time.sleep(0.02)
            per_call = []
            t0 = time.perf_counter()
            for cwn, sn in names:
                t1 = time.perf_counter()
                rr = subprocess.run([LDPC, "0", "0", "0", "7", str(a.max_iter), "1", cwn, sn, "decode_n18432_m2048_final", "0", "0", "0", "0"],
                                    cwd=dirs["worker"], capture_output=True, env=env)
                per_call.append(time.perf_counter() - t1)
                if rr.returncode != 0:
                    raise SystemExit("ldpc client failed: " + rr.stderr.decode())
            out["worker_serial_s"] = time.perf_counter() - t0
            out["worker_first_call_s"] = per_call[0]
            out["worker_median_call_ms"] = 1e3 * float(np.median(per_call[1:])) if len(per_call) > 1 else None
        finally:
            subprocess.run([LDPC, "--shutdown"], env=env, capture_output=True)
            worker.wait(timeout=60)
        # reference: one process per codeword, serially (os.system loop of decoder.py)
        nref = min(a.ref_frames, a.frames)
        t0 = time.perf_counter()
        for cwn, sn in names[:nref]:
            rr = subprocess.run([REF, "0", "0", "0", "7", str(a.max_iter), "1", cwn, sn, "decode_n18432_m2048_final", "0", "0", "0", "0"],
                                cwd=dirs["ref"], capture_output=True)
            if rr.returncode != 0:
                raise SystemExit("ldpc_ref failed")
        dt = time.perf_counter() - t0
        out["reference_processes_run"] = nref
        out["reference_serial_s"] = dt * a.frames / nref
        same = all(open(os.path.join(dirs["ours"], "dec_%s.txt" % cwn), "rb").read() == open(os.path.join(dirs["ref"], "dec_%s.txt" % cwn), "rb").read()
                   for cwn, _ in names[:nref])
        same_w = all(open(os.path.join(dirs["worker"], "dec_%s.txt" % cwn), "rb").read() == open(os.path.join(dirs["ref"], "dec_%s.txt" % cwn), "rb").read()
                     for cwn, _ in names[:nref])
        out["dec_files_identical"] = bool(same)
        out["worker_dec_files_identical"] = bool(same_w)
        out["speedup_process_level"] = out["reference_serial_s"] / out["ours_one_process_s"]
        out["speedup_unmodified_pipeline"] = out["reference_serial_s"] / out["worker_serial_s"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
