"""In-order issue model of a SASS loop body (lone warp): an offline aid for scheduling experiments on the compositor's blend loop.

    python ubench/sass_model.py <libgsr.so> <kernel-name-substring> [--lat-packed 9]

Extracts the innermost loop that contains packed f32x2 instructions from `cuobjdump -sass`, then replays it for a few iterations
with: one issue per cycle, fixed-latency pipes (FMA/ALU: result ready `lat` cycles after issue, reciprocal throughput 2 cycles per
pipe per SMSP, B300_MICROARCH.md "Pipe rates & latencies"), packed f32x2 ops with their own (longer) dependent-issue latency, and
shared-memory loads at ~30 cycles.  Output: cycles per loop iteration for ONE warp alone on its SMSP and the FMA-pipe bound.
Calibration: the round-1 kernel's loop (171 instructions per 4 splats) measured 522 cycles per iteration on a B200 (17 us per
256-splat chunk, schedule trace); --lat-packed is the knob that reproduces it."""
import re
import subprocess
import sys

FMA = {"FFMA2", "FMUL2", "FADD2", "FFMA", "FMUL", "FADD", "IMAD", "HFMA2"}
ALU = {"FMNMX", "FSETP", "FSEL", "LEA", "IADD3", "LOP3", "SEL", "ISETP", "MOV", "SHF", "PRMT", "VIADD", "VIMNMX", "I2FP", "FMNMX3"}
XU = {"F2I", "I2F", "MUFU", "F2F"}


def kernel_sass(lib, name):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    blocks = out.split("Function : ")
    for b in blocks[1:]:
        head = b.split("\n", 1)[0]
        if name in head:
            return head, b
    raise SystemExit(f"no function matching {name!r}")


def parse(body):
    ins = []
    for l in body.splitlines():
        m = re.search(r"/\*([0-9a-f]{4})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    return ins


def find_loop(ins):
    packed = [i for i, (_, t) in enumerate(ins) if re.search(r"\b(FFMA2|FMUL2|FADD2)\b", t)]
    best = None
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA\s+(0x[0-9a-f]+)", t)
        if not m:
            continue
        tgt = int(m.group(1), 16)
        if tgt < a:  # backward branch
            lo = next(k for k, (aa, _) in enumerate(ins) if aa >= tgt)
            n = sum(1 for p in packed if lo <= p <= i)
            if n and (best is None or (i - lo) < (best[1] - best[0])):
                best = (lo, i, n)
    if best is None:
        raise SystemExit("no loop with packed instructions found")
    return ins[best[0]:best[1] + 1]


def regs_of(tok, wide_default=1):
    """registers named by one operand token, e.g. R52.F32x2.HI_LO -> [52, 53]; -R72.F32x2.HI_LO; R54.F32 -> [54]; R60.reuse..."""
    m = re.match(r"[-!|~]*R(\d+)(.*)", tok)
    if not m:
        return []
    r = int(m.group(1))
    suf = m.group(2)
    n = 2 if "F32x2" in suf or ".64" in suf else wide_default
    return [r + k for k in range(n)]


def decode(text):
    pred = None
    t = text
    m = re.match(r"@(!?)(U?P\d+)\s+(.*)", t)
    if m:
        pred = m.group(2)
        t = m.group(3)
    op_full, _, rest = t.partition(" ")
    op = op_full.split(".")[0]
    ops = [o.strip() for o in rest.split(",")] if rest else []
    dst, src = [], []
    width = 1
    if op in ("FFMA2", "FMUL2", "FADD2"):
        width = 2
    if op == "LDS":
        width = 4 if ".128" in op_full else (2 if ".64" in op_full else 1)
    if op in ("BRA", "BAR", "VOTE", "VOTEU", "UIADD3", "ULEA", "UMOV", "BSSY", "BSYNC", "NOP"):
        # predicate reads/writes still matter for VOTE/BRA
        preds_w = [o for o in ops[:1] if re.match(r"P\d+$", o)] if op.startswith("VOTE") else []
        preds_r = [p for p in re.findall(r"P\d+", rest)] if op == "BRA" or op.startswith("VOTE") else []
        return dict(op=op, dst=[("P", p) for p in preds_w], src=[("P", p) for p in preds_r] + ([("P", pred.lstrip('!'))] if pred else []), pipe="ctl")
    if ops:
        d = ops[0]
        if re.match(r"P\d+$", d):  # FSETP / ISETP: P0, PT, a, b, PT
            dst = [("P", d)]
            for o in ops[1:]:
                if re.match(r"!?P\d+$", o):
                    src.append(("P", o.lstrip("!")))
                src += [("R", r) for r in regs_of(o)]
        else:
            dst = [("R", r) for r in regs_of(d, width)] if op != "LDS" else [("R", int(re.match(r"R(\d+)", d).group(1)) + k) for k in range(width)]
            for o in ops[1:]:
                if re.match(r"!?P\d+$", o):
                    src.append(("P", o.lstrip("!")))
                else:
                    src += [("R", r) for r in regs_of(o)]
    if pred:
        src.append(("P", pred.lstrip("!")))
    pipe = "fma" if op in FMA else "alu" if op in ALU else "xu" if op in XU else "lds" if op == "LDS" else "other"
    return dict(op=op, dst=dst, src=src, pipe=pipe, packed=op in ("FFMA2", "FMUL2", "FADD2"))


def simulate(loop, lat_packed=9, lat_fma=4, lat_alu=4, lat_lds=30, lat_xu=14, iters=6):
    dec = [decode(t) for _, t in loop]
    ready = {}
    pipe_free = {"fma": 0, "alu": 0, "xu": 0, "lds": 0}
    rt = {"fma": 2, "alu": 2, "xu": 8, "lds": 2}
    cyc = 0
    marks = []
    for it in range(iters):
        marks.append(cyc)
        for d in dec:
            t = cyc + 1
            for s in d["src"]:
                t = max(t, ready.get(s, 0))
            p = d["pipe"]
            if p in pipe_free:
                t = max(t, pipe_free[p])
                pipe_free[p] = t + rt[p]
            lat = lat_packed if d.get("packed") else {"fma": lat_fma, "alu": lat_alu, "xu": lat_xu, "lds": lat_lds}.get(p, 2)
            for w in d["dst"]:
                ready[w] = t + lat
            cyc = t
    per = (marks[-1] - marks[1]) / (len(marks) - 2)
    n_fma = sum(1 for d in dec if d["pipe"] == "fma")
    n_alu = sum(1 for d in dec if d["pipe"] == "alu")
    return per, len(dec), n_fma, n_alu


if __name__ == "__main__":
    lib, name = sys.argv[1], sys.argv[2]
    lp = 9
    if "--lat-packed" in sys.argv:
        lp = int(sys.argv[sys.argv.index("--lat-packed") + 1])
    head, body = kernel_sass(lib, name)
    loop = find_loop(parse(body))
    for l in (6, 7, 8, 9, 10, 11):
        per, n, nf, na = simulate(loop, lat_packed=l)
        print(f"{head[:60]}: {n} instr/iter ({nf} FMA-pipe, {na} ALU-pipe) lat_packed={l}: {per:.0f} cycles/iter alone; FMA-pipe bound {2 * nf}")
