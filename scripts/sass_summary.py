"""profiles/r2_sass_tc.txt: opcode evidence that the field kernels are Blackwell-native -- histogram of the tcgen05 / TMEM /
TMA / cluster opcodes per kernel in libnsb.so, plus the MMA issue loop of the inference forward kernel.  CPU only (cuobjdump)."""
import collections, os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, "nerf_sandbox_b200", "libnsb.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, body = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        body[kern] = []
    elif kern and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
        body[kern].append(line)
want = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "UTCCP", "SYNCS", "UCGABAR", "LDGMC", "F2FP", "MUFU", "REDG", "ATOMG")
print(f"# {os.path.relpath(lib, root)} -- cuobjdump -sass, sm_100a\n")
print("## tcgen05 / TMEM / TMA / cluster opcodes per kernel (static counts)\n")
for k, lines in body.items():
    ops = collections.Counter()
    for l in lines:
        t = l.split("*/", 1)[1].split()
        op = t[1] if t[0].startswith("@") else t[0]
        for w in want:
            if op.startswith(w):
                ops[".".join(op.rstrip(";").split(".")[:4])] += 1
    if any(o.startswith(("UTCHMMA", "LDTM", "UTMALDG", "UBLKCP", "LDGMC")) for o in ops):
        print(f"{k}  ({len(lines)} instructions)")
        for o, n in sorted(ops.items()):
            print(f"    {n:5d}  {o}")
print("\n## MMA issue loop of field_fwd_kernel<false,false> (inference, fp16 operands): descriptor updates + UTCHMMA.2CTA + commits\n")
for k, lines in body.items():
    if "field_fwd_kernel<false, false>" in k or "field_fwd_kernel<(bool)0, (bool)0>" in k:
        idx = [i for i, l in enumerate(lines) if "UTCHMMA" in l]
        if idx:
            for l in lines[max(0, idx[0] - 14): idx[min(3, len(idx) - 1)] + 10]:
                print(l.rstrip())
        break
