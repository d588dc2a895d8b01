"""TEST INFRASTRUCTURE — builds tests/_build/libodis_b200_emu.so: the library's own sources (csrc/*.cu, engine included) with
the CUDA-only syntax rewritten for a host compiler and compiled against tests/simt/simt_emu.h (a host emulation of the CUDA
execution model). Host .cpp files are compiled as they are. Used by tests/test_emulated_kernels.py to run kernel logic and the
engine's launch sequencing on the CPU; it is never loaded by the product path (geodesicodis_b200/_lib.py loads
libodis_b200.so) and is no substitute for the B200 runs.

The bulk-async staged kernels run too: mbarrier objects, cp.async.bulk (completing at once) and named barriers are modelled in
simt_emu.h. The FP64 tensor-core mma.sync.m8n8k4 of the ensemble GEMMs is modelled by its fragment layout. Streams are worker threads, so
several solvers ("devices") of one process run concurrently and exchange halos through each other's memory and flags."""
from __future__ import annotations

import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "geodesicodis_b200", "csrc")
OUT = os.path.join(ROOT, "tests", "_build")
LIB = os.path.join(OUT, "libodis_b200_emu.so")

CUDA_SOURCES = ["odis_kernels.cu", "odis_kernels_pipe.cu", "odis_kernels_nl.cu", "odis_sh.cu", "odis_ensemble.cu", "odis_engine.cu"]
SWITCH_SRC = r'''// context switch of the SIMT emulation's fibers (tests/simt/simt_emu.h): x86-64 System V, callee-saved registers + stack pointer
#if !defined(__x86_64__)
#error "the SIMT emulation's context switch is written for x86-64"
#endif
asm(R"(
    .pushsection .text
    .globl simt_switch
    .type simt_switch,@function
simt_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
    .size simt_switch,.-simt_switch
    .popsection
)");
'''


def _match_paren(s: str, i: int) -> int:
    """index of the ')' matching the '(' at s[i] (string literals skipped)"""
    depth, k = 0, i
    while k < len(s):
        c = s[k]
        if c == '"':
            k += 1
            while s[k] != '"':
                k += 2 if s[k] == "\\" else 1
        elif c == "(":
            depth += 1
        elif c == ")":
            depth -= 1
            if depth == 0:
                return k
        k += 1
    raise ValueError("unbalanced parentheses")


def _split_top(s: str, sep: str) -> list[str]:
    parts, depth, cur, k = [], 0, [], 0
    while k < len(s):
        c = s[k]
        if c == '"':
            j = k + 1
            while s[j] != '"':
                j += 2 if s[j] == "\\" else 1
            cur.append(s[k:j + 1]); k = j + 1
            continue
        if c in "([{":
            depth += 1
        elif c in ")]}":
            depth -= 1
        if c == sep and depth == 0:
            parts.append("".join(cur)); cur = []
        else:
            cur.append(c)
        k += 1
    parts.append("".join(cur))
    return parts


def rewrite_launches(s: str) -> tuple[str, int]:
    n = 0
    while True:
        p = s.find("<<<")
        if p < 0:
            return s, n
        i = p
        if s[i - 1] == ">":                                 # template arguments of the kernel
            depth = 0
            while True:
                i -= 1
                if s[i] == ">":
                    depth += 1
                elif s[i] == "<":
                    depth -= 1
                    if depth == 0:
                        break
        while i > 0 and (s[i - 1].isalnum() or s[i - 1] in "_:"):
            i -= 1
        kernel = s[i:p]
        q = s.index(">>>", p)
        cfg = [c.strip() for c in _split_top(s[p + 3:q], ",")]
        cfg += ["0", "nullptr"][len(cfg) - 2:] if len(cfg) < 4 else []
        a = q + 3
        while s[a].isspace():
            a += 1
        assert s[a] == "(", s[p - 40:q + 40]
        b = _match_paren(s, a)
        args = s[a + 1:b].strip()
        s = s[:i] + f"simt::launch({cfg[0]}, {cfg[1]}, {cfg[2]}, {cfg[3]}, {kernel}{', ' + args if args else ''})" + s[b + 1:]
        n += 1


def _operands(section: str) -> list[str]:
    out = []
    for part in _split_top(section, ","):
        part = part.strip()
        if not part:
            continue
        a = part.index("(")
        out.append(part[a + 1:_match_paren(part, a)].strip())
    return out


def rewrite_asm(s: str) -> tuple[str, int]:
    n = 0
    for m in reversed(list(re.finditer(r"\basm\s*(?:volatile\s*)?\(", s))):
        a = m.end() - 1
        b = _match_paren(s, a)
        end = b + 1
        while s[end].isspace():
            end += 1
        assert s[end] == ";", s[m.start():end + 10]
        sections = _split_top(s[a + 1:b], ":")
        ptx = "".join(re.findall(r'"((?:[^"\\]|\\.)*)"', sections[0]))
        outs = _operands(sections[1]) if len(sections) > 1 else []
        ins = _operands(sections[2]) if len(sections) > 2 else []
        if ptx.startswith("createpolicy") or "%globaltimer" in ptx:      # (the latter: tracing aid of the variant library, compiled out here)
            rep = f"{outs[0]} = 0;"
        elif ptx.startswith("st.volatile.global.v4.u32"):        # LL line: each 8-byte half {data, tag} lands atomically
            rep = (f"{{ unsigned long long* simt_p_ = (unsigned long long*)({ins[0]}); "
                   f"__atomic_store_n(simt_p_, (unsigned long long)({ins[1]}) | ((unsigned long long)({ins[2]}) << 32), __ATOMIC_RELEASE); "
                   f"__atomic_store_n(simt_p_ + 1, (unsigned long long)({ins[3]}) | ((unsigned long long)({ins[4]}) << 32), __ATOMIC_RELEASE); }}")
        elif ptx.startswith("ld.volatile.global.v4.u32"):
            rep = (f"{{ const unsigned long long* simt_p_ = (const unsigned long long*)({ins[0]}); std::this_thread::yield(); "
                   f"const unsigned long long simt_a_ = __atomic_load_n(simt_p_, __ATOMIC_ACQUIRE), simt_b_ = __atomic_load_n(simt_p_ + 1, __ATOMIC_ACQUIRE); "
                   f"{outs[0]} = (unsigned int)simt_a_; {outs[1]} = (unsigned int)(simt_a_ >> 32); {outs[2]} = (unsigned int)simt_b_; {outs[3]} = (unsigned int)(simt_b_ >> 32); }}")
        elif ptx.startswith("ld.") and "{%0, %1}" in ptx:
            rep = f"{{ const auto* simt_p_ = ({ins[0]}); {outs[0]} = simt_p_->x; {outs[1]} = simt_p_->y; }}"
        elif ptx.startswith("cp.async.bulk.prefetch"):          # no data moves; touching both ends lets the address sanitizer check the range
            rep = f"{{ const volatile unsigned char* simt_p_ = (const volatile unsigned char*)({ins[0]}); (void)simt_p_[0]; (void)simt_p_[({ins[1]}) - 1]; }}"
        elif ptx.startswith("ld.acquire"):
            rep = f"{outs[0]} = (std::this_thread::yield(), __atomic_load_n({ins[0]}, __ATOMIC_ACQUIRE));"     # flag polls: let the other "devices" run
        elif ptx.startswith("ld.relaxed.sys"):
            rep = f"{outs[0]} = *(const volatile decltype({outs[0]})*)({ins[0]});"
        elif ptx.startswith("ld."):
            rep = f"{outs[0]} = *({ins[0]});"
        elif ptx.startswith("st.release") or ptx.startswith("st.relaxed.sys"):     # (a relaxed flag store behind a fence: modelled as a release store)
            rep = f"__atomic_store_n({ins[0]}, (unsigned long long)({ins[1]}), __ATOMIC_RELEASE);"
        elif ptx.startswith("st."):
            rep = f"*({ins[0]}) = ({ins[1]});"
        elif ptx.startswith("mbarrier.init"):
            rep = f"simt::mbar_init({ins[0]}, {ins[1]});"
        elif ptx.startswith("mbarrier.arrive.expect_tx"):
            rep = f"simt::mbar_arrive({ins[0]}, {ins[1]});"
        elif ptx.startswith("mbarrier.arrive"):
            rep = f"simt::mbar_arrive({ins[0]}, 0);"
        elif "mbarrier.try_wait.parity" in ptx:
            rep = f"simt::mbar_wait({ins[0]}, {ins[1]});"
        elif ptx.startswith("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes"):
            rep = f"simt::bulk_copy({ins[0]}, {ins[1]}, {ins[2]}, {ins[3]});"
        elif ptx.startswith("fence.") or ptx.startswith("griddepcontrol."):
            rep = ";"
        elif ptx.startswith("bar.sync %0, %1"):
            rep = f"simt::named_barrier({ins[0]}, {ins[1]});"
        elif ptx.startswith("bar.sync 1, %0"):
            rep = f"simt::named_barrier(1, {ins[0]});"
        elif ptx.startswith("mma.sync.aligned.m8n8k4.row.col.f64"):
            rep = f"simt::dmma_m8n8k4({outs[0]}, {outs[1]}, {ins[0]}, {ins[1]});"

        else:
            raise ValueError("no emulation for PTX: " + ptx)
        s = s[:m.start()] + rep + s[end + 1:]
        n += 1
    return s, n


def transform(name: str, text: str) -> str:
    text, n_launch = rewrite_launches(text)
    text, n_asm = rewrite_asm(text)
    text, n_dyn = re.subn(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?((?:unsigned\s+)?\w+)\s+(\w+)\s*\[\s*\]\s*;",
                          r"\1* \2 = reinterpret_cast<\1*>(simt::dynamic_shared());", text)
    return f"// generated by tests/simt/build_emu.py from csrc/{name}: {n_launch} launches, {n_asm} asm statements, {n_dyn} dynamic shared arrays rewritten\n" + text


def build(force: bool = False, verbose: bool = False, asan: bool = False, tsan: bool = False) -> str:
    """asan: AddressSanitizer build (libodis_b200_emu_asan.so; load with LD_PRELOAD=$(gcc -print-file-name=libasan.so)): device
    arrays are host heap blocks, so an out-of-range access in a kernel is reported like compute-sanitizer's memcheck would.
    tsan: ThreadSanitizer build (libodis_b200_emu_tsan.so, LD_PRELOAD libtsan.so, OMP_NUM_THREADS=1): every stream is a thread, so an
    access of one "device" that is not ordered after another device's (or the host's) write by a flag, an event or a stream
    synchronisation is reported as a data race -- a racecheck of the halo / all-reduce protocol and of the engine's host-side waits.
    Function entry/exit instrumentation is off because the fibers switch stacks behind the sanitizer's back."""
    from geodesicodis_b200.build import HOST_SOURCES, HOST_FLAGS
    global LIB
    tag = "_asan" if asan else "_tsan" if tsan else ""
    lib = os.path.join(OUT, f"libodis_b200_emu{tag}.so")
    LIB = lib
    src_dir = os.path.join(OUT, "emu_src" + tag)
    os.makedirs(src_dir, exist_ok=True)
    newest = max(os.path.getmtime(os.path.join(CSRC, f)) for f in os.listdir(CSRC))
    newest = max(newest, os.path.getmtime(os.path.join(HERE, "simt_emu.h")), os.path.getmtime(os.path.abspath(__file__)),
                 os.path.getmtime(os.path.join(ROOT, "include", "odis_b200.h")))
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= newest:
        return LIB
    san = ["-g", "-fsanitize=address", "-fno-omit-frame-pointer"] if asan else []
    if tsan:
        san = ["-g", "-fsanitize=thread", "--param=tsan-instrument-func-entry-exit=0"]
    flags = ["-O1", *san, "-mtls-dialect=gnu2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-w", "-I" + os.path.join(HERE, "stub"), "-I" + CSRC]
    jobs = []
    for name in sorted(os.listdir(CSRC)):                       # device headers with inline PTX get the same rewrite: the transformed
        if name.endswith(".cuh"):                               # copy sits beside the transformed sources, where #include "..." looks first
            with open(os.path.join(CSRC, name)) as f:
                text = transform(name, f.read())
            with open(os.path.join(src_dir, name), "w") as f:
                f.write(text)
    for name in CUDA_SOURCES:
        with open(os.path.join(CSRC, name)) as f:
            text = transform(name, f.read())
        text = text.replace('#include "../../include/odis_b200.h"', f'#include "{os.path.join(ROOT, "include", "odis_b200.h")}"')
        path = os.path.join(src_dir, name + ".cpp")
        with open(path, "w") as f:
            f.write(text)
        jobs.append((["g++", *flags, "-c", path, "-o", path + ".o"], path + ".o"))
    switch = os.path.join(src_dir, "simt_switch.cpp")
    with open(switch, "w") as f:
        f.write(SWITCH_SRC)
    jobs.append((["g++", *flags, "-c", switch, "-o", switch + ".o"], switch + ".o"))
    for name in HOST_SOURCES:                                   # host code as it is (no CUDA in it)
        obj = os.path.join(src_dir, name + ".o")
        jobs.append((["g++", *HOST_FLAGS, "-w", "-c", os.path.join(CSRC, name), "-o", obj], obj))
    procs = [(subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True), cmd) for cmd, _ in jobs]
    for p, cmd in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("emulation build failed: %s\n%s" % (" ".join(cmd), out[-6000:]))
        if verbose and out.strip():
            print(out)
    link = ["g++", "-shared", *(["-fsanitize=address"] if asan else ["-fsanitize=thread"] if tsan else []), "-o", LIB, *[o for _, o in jobs], "-fopenmp", "-pthread"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("emulation link failed:\n" + r.stdout[-4000:])
    return LIB


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, asan="--asan" in sys.argv, tsan="--tsan" in sys.argv))
