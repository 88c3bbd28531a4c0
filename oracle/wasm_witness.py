"""Minimal WebAssembly interpreter + circom-2 witness-calculator host protocol (TEST INFRASTRUCTURE ONLY).

The reference computes witnesses by running the circom-generated .wasm under wasmer
(/root/reference/ark-circom/src/witness/witness_calculator.rs:56-255); no WASM runtime exists in this image, so
`tests/golden/make_golden.py` uses this integer-only interpreter once to produce the sha256 witness fixture
(fixture F2 of SURVEY 8c).  Supports exactly what circom 2.1 emits: i32/i64 arithmetic, loads/stores, structured
control flow, direct calls, one exported memory, data segments; no floats, no tables.

Host protocol (witness_calculator.rs:219-255): init(sanity) ; n32 = getFieldNumLen32() ; per input value: write its
n32 little-endian u32 limbs with writeSharedRWMemory(j, limb) then setInputSignal(hash_msb, hash_lsb, index) with the
FNV-1a-64 hash of the signal name (witness/mod.rs:18-24) ; getWitnessSize() ; per index getWitness(i) + n32 x
readSharedRWMemory(j).
"""
from __future__ import annotations

M32, M64 = 0xFFFFFFFF, 0xFFFFFFFFFFFFFFFF


def _leb_u(b, p):
    r = s = 0
    while True:
        x = b[p]
        p += 1
        r |= (x & 0x7F) << s
        s += 7
        if not x & 0x80:
            return r, p


def _leb_s(b, p, bits):
    r = s = 0
    while True:
        x = b[p]
        p += 1
        r |= (x & 0x7F) << s
        s += 7
        if not x & 0x80:
            if x & 0x40:
                r -= 1 << s
            return r & ((1 << bits) - 1), p


class Module:
    def __init__(self, data: bytes):
        assert data[:8] == b"\0asm\1\0\0\0"
        self.types, self.imports, self.func_types, self.exports, self.codes = [], [], [], {}, []
        self.mem_pages = 0
        self.globals = []
        self.data_segs = []
        p = 8
        while p < len(data):
            sid = data[p]
            size, p = _leb_u(data, p + 1)
            end = p + size
            if sid == 1:
                n, p = _leb_u(data, p)
                for _ in range(n):
                    assert data[p] == 0x60
                    np_, p = _leb_u(data, p + 1)
                    p += np_
                    nr, p = _leb_u(data, p)
                    p += nr
                    self.types.append((np_, nr))
            elif sid == 2:
                n, p = _leb_u(data, p)
                for _ in range(n):
                    ml, p = _leb_u(data, p)
                    mod = data[p:p + ml].decode(); p += ml
                    nl, p = _leb_u(data, p)
                    name = data[p:p + nl].decode(); p += nl
                    kind = data[p]; p += 1
                    if kind == 0:
                        ti, p = _leb_u(data, p)
                        self.imports.append((mod, name, ti))
                    elif kind == 2:
                        fl = data[p]; p += 1
                        mn, p = _leb_u(data, p)
                        if fl & 1:
                            _, p = _leb_u(data, p)
                        self.mem_pages = mn
                    else:
                        raise NotImplementedError("import kind %d" % kind)
            elif sid == 3:
                n, p = _leb_u(data, p)
                for _ in range(n):
                    ti, p = _leb_u(data, p)
                    self.func_types.append(ti)
            elif sid == 5:
                n, p = _leb_u(data, p)
                fl = data[p]; p += 1
                mn, p = _leb_u(data, p)
                if fl & 1:
                    _, p = _leb_u(data, p)
                self.mem_pages = mn
            elif sid == 6:
                n, p = _leb_u(data, p)
                for _ in range(n):
                    vt, mut = data[p], data[p + 1]; p += 2
                    op = data[p]; p += 1
                    if op == 0x41:
                        v, p = _leb_s(data, p, 32)
                    elif op == 0x42:
                        v, p = _leb_s(data, p, 64)
                    else:
                        raise NotImplementedError
                    assert data[p] == 0x0B; p += 1
                    self.globals.append(v)
            elif sid == 7:
                n, p = _leb_u(data, p)
                for _ in range(n):
                    nl, p = _leb_u(data, p)
                    name = data[p:p + nl].decode(); p += nl
                    kind = data[p]; p += 1
                    idx, p = _leb_u(data, p)
                    if kind == 0:
                        self.exports[name] = idx
            elif sid == 10:
                n, p = _leb_u(data, p)
                for _ in range(n):
                    bs, p = _leb_u(data, p)
                    body_end = p + bs
                    nl, q = _leb_u(data, p)
                    nlocals = 0
                    for _ in range(nl):
                        cnt, q = _leb_u(data, q)
                        q += 1
                        nlocals += cnt
                    self.codes.append((nlocals, data[q:body_end]))
                    p = body_end
            elif sid == 11:
                n, p = _leb_u(data, p)
                for _ in range(n):
                    fl, p = _leb_u(data, p)
                    assert fl == 0
                    assert data[p] == 0x41
                    off, p = _leb_s(data, p + 1, 32)
                    assert data[p] == 0x0B; p += 1
                    ln, p = _leb_u(data, p)
                    self.data_segs.append((off, data[p:p + ln])); p += ln
            p = end
        self.n_imports = len(self.imports)
        self.decoded = [None] * len(self.codes)

    # -- decode a function body into (ops, imms) with resolved branch targets --------------------
    def decode(self, fi):
        if self.decoded[fi] is not None:
            return self.decoded[fi]
        nlocals, code = self.codes[fi]
        ops, imm = [], []
        stack = []          # (kind, index_of_block_op)
        p = 0
        n = len(code)
        while p < n:
            op = code[p]; p += 1
            a = 0
            if op in (0x02, 0x03, 0x04):          # block loop if (blocktype)
                bt = code[p]; p += 1
                arity = 0 if bt == 0x40 else 1
                a = [arity, None, None]            # arity, end_index, else_index
                stack.append(len(ops))
            elif op == 0x05:
                ops_i = stack[-1]
                imm[ops_i][2] = len(ops)
            elif op == 0x0B:
                if stack:
                    ops_i = stack.pop()
                    imm[ops_i][1] = len(ops)
            elif op in (0x0C, 0x0D, 0x10, 0x20, 0x21, 0x22, 0x23, 0x24):
                a, p = _leb_u(code, p)
            elif 0x28 <= op <= 0x3E:
                _, p = _leb_u(code, p)
                a, p = _leb_u(code, p)
            elif op in (0x3F, 0x40):
                p += 1
            elif op == 0x41:
                a, p = _leb_s(code, p, 32)
            elif op == 0x42:
                a, p = _leb_s(code, p, 64)
            ops.append(op); imm.append(a)
        self.decoded[fi] = (nlocals, ops, imm)
        return self.decoded[fi]


class Instance:
    def __init__(self, mod: Module, host):
        self.m = mod
        self.mem = bytearray(max(mod.mem_pages, 1) * 65536)
        for off, seg in mod.data_segs:
            self.mem[off:off + len(seg)] = seg
        self.globals = list(mod.globals)
        self.host = host
        self.steps = 0

    def call(self, name, *args):
        return self.invoke(self.m.exports[name], list(args))

    def invoke(self, fidx, args):
        m = self.m
        if fidx < m.n_imports:
            return self.host(m.imports[fidx][1], args)
        fi = fidx - m.n_imports
        nparams, nres = m.types[m.func_types[fi]]
        nlocals, ops, imm = m.decode(fi)
        loc = args + [0] * nlocals
        st = []
        labels = []      # (target_pc_for_br, stack_height, arity, is_loop)
        mem = self.mem
        pc = 0
        nops = len(ops)
        while pc < nops:
            op = ops[pc]; a = imm[pc]; pc += 1
            if op == 0x20: st.append(loc[a])
            elif op == 0x21: loc[a] = st.pop()
            elif op == 0x22: loc[a] = st[-1]
            elif op == 0x41 or op == 0x42: st.append(a)
            elif op == 0x6A: b = st.pop(); st[-1] = (st[-1] + b) & M32
            elif op == 0x7C: b = st.pop(); st[-1] = (st[-1] + b) & M64
            elif op == 0x7E: b = st.pop(); st[-1] = (st[-1] * b) & M64
            elif op == 0x29: ad = st.pop() + a; st.append(int.from_bytes(mem[ad:ad + 8], "little"))
            elif op == 0x28: ad = st.pop() + a; st.append(int.from_bytes(mem[ad:ad + 4], "little"))
            elif op == 0x35: ad = st.pop() + a; st.append(int.from_bytes(mem[ad:ad + 4], "little"))
            elif op == 0x37: v = st.pop(); ad = st.pop() + a; mem[ad:ad + 8] = v.to_bytes(8, "little")
            elif op == 0x36: v = st.pop(); ad = st.pop() + a; mem[ad:ad + 4] = (v & M32).to_bytes(4, "little")
            elif op == 0x3E: v = st.pop(); ad = st.pop() + a; mem[ad:ad + 4] = (v & M32).to_bytes(4, "little")
            elif op == 0x88: b = st.pop(); st[-1] = st[-1] >> (b & 63)
            elif op == 0x86: b = st.pop(); st[-1] = (st[-1] << (b & 63)) & M64
            elif op == 0x83: b = st.pop(); st[-1] &= b
            elif op == 0x84: b = st.pop(); st[-1] |= b
            elif op == 0xAD: pass                                  # i64.extend_i32_u
            elif op == 0xA7: st[-1] &= M32                         # i32.wrap_i64
            elif op == 0x02: labels.append((a[1], len(st), a[0], False))
            elif op == 0x03: labels.append((pc - 1, len(st), 0, True))
            elif op == 0x04:
                c = st.pop()
                labels.append((a[1], len(st), a[0], False))
                if not c:
                    if a[2] is not None:
                        pc = a[2] + 1
                    else:
                        pc = a[1]          # jump to the end op (pops the label)
            elif op == 0x05:                                       # else reached from the then-branch: skip to end
                pc = labels[-1][0]
            elif op == 0x0B:
                if labels:
                    labels.pop()
            elif op == 0x0C or op == 0x0D:
                if op == 0x0D and not st.pop():
                    continue
                if a >= len(labels):                                # branch to the function label == return
                    break
                tgt, h, ar, is_loop = labels[-1 - a]
                if is_loop:
                    del labels[len(labels) - a:]
                    del st[h:]
                    pc = tgt + 1                                   # re-enter just after the loop op (label kept)
                else:
                    res = st[len(st) - ar:] if ar else []
                    del st[h:]
                    st.extend(res)
                    del labels[len(labels) - 1 - a + 1:]
                    pc = tgt                                       # the end op pops the target label
            elif op == 0x0F:
                break
            elif op == 0x10:
                if a < m.n_imports:
                    np_, nr = m.types[m.imports[a][2]]
                else:
                    np_, nr = m.types[m.func_types[a - m.n_imports]]
                cargs = st[len(st) - np_:] if np_ else []
                if np_:
                    del st[len(st) - np_:]
                r = self.invoke(a, cargs)
                if nr:
                    st.append(r)
            elif op == 0x1A: st.pop()
            elif op == 0x1B: c = st.pop(); b = st.pop(); st[-1] = st[-1] if c else b
            elif op == 0x23: st.append(self.globals[a])
            elif op == 0x24: self.globals[a] = st.pop()
            elif op == 0x2D: ad = st.pop() + a; st.append(mem[ad])
            elif op == 0x31: ad = st.pop() + a; st.append(mem[ad])
            elif op == 0x3A or op == 0x3C: v = st.pop(); ad = st.pop() + a; mem[ad] = v & 0xFF
            elif op == 0x2F or op == 0x33: ad = st.pop() + a; st.append(int.from_bytes(mem[ad:ad + 2], "little"))
            elif op == 0x3B or op == 0x3D: v = st.pop(); ad = st.pop() + a; mem[ad:ad + 2] = (v & 0xFFFF).to_bytes(2, "little")
            elif op == 0x34:
                ad = st.pop() + a; v = int.from_bytes(mem[ad:ad + 4], "little")
                st.append(v | (M64 ^ M32) if v & 0x80000000 else v)
            elif op == 0x3F: st.append(len(mem) // 65536)
            elif op == 0x40:
                n = st.pop(); old = len(mem) // 65536
                mem.extend(bytes(n * 65536)); st.append(old)
            elif op == 0x45: st[-1] = 1 if st[-1] == 0 else 0
            elif op == 0x50: st[-1] = 1 if st[-1] == 0 else 0
            elif op == 0x46 or op == 0x51: b = st.pop(); st[-1] = 1 if st[-1] == b else 0
            elif op == 0x47 or op == 0x52: b = st.pop(); st[-1] = 1 if st[-1] != b else 0
            elif op == 0x49 or op == 0x54: b = st.pop(); st[-1] = 1 if st[-1] < b else 0
            elif op == 0x4B or op == 0x56: b = st.pop(); st[-1] = 1 if st[-1] > b else 0
            elif op == 0x4D or op == 0x58: b = st.pop(); st[-1] = 1 if st[-1] <= b else 0
            elif op == 0x4F or op == 0x5A: b = st.pop(); st[-1] = 1 if st[-1] >= b else 0
            elif op in (0x48, 0x4A, 0x4C, 0x4E):                   # i32 signed compares
                b = st.pop(); x = st[-1]
                x = x - (1 << 32) if x & 0x80000000 else x
                b = b - (1 << 32) if b & 0x80000000 else b
                st[-1] = int({0x48: x < b, 0x4A: x > b, 0x4C: x <= b, 0x4E: x >= b}[op])
            elif op in (0x53, 0x55, 0x57, 0x59):                   # i64 signed compares
                b = st.pop(); x = st[-1]
                x = x - (1 << 64) if x >> 63 else x
                b = b - (1 << 64) if b >> 63 else b
                st[-1] = int({0x53: x < b, 0x55: x > b, 0x57: x <= b, 0x59: x >= b}[op])
            elif op == 0x6B: b = st.pop(); st[-1] = (st[-1] - b) & M32
            elif op == 0x6C: b = st.pop(); st[-1] = (st[-1] * b) & M32
            elif op == 0x6E: b = st.pop(); st[-1] = st[-1] // b
            elif op == 0x70: b = st.pop(); st[-1] = st[-1] % b
            elif op == 0x71: b = st.pop(); st[-1] &= b
            elif op == 0x72: b = st.pop(); st[-1] |= b
            elif op == 0x73: b = st.pop(); st[-1] ^= b
            elif op == 0x74: b = st.pop(); st[-1] = (st[-1] << (b & 31)) & M32
            elif op == 0x76: b = st.pop(); st[-1] = st[-1] >> (b & 31)
            elif op == 0x75:
                b = st.pop() & 31; x = st[-1]
                x = x - (1 << 32) if x & 0x80000000 else x
                st[-1] = (x >> b) & M32
            elif op == 0x7D: b = st.pop(); st[-1] = (st[-1] - b) & M64
            elif op == 0x80: b = st.pop(); st[-1] = st[-1] // b
            elif op == 0x82: b = st.pop(); st[-1] = st[-1] % b
            elif op == 0x85: b = st.pop(); st[-1] ^= b
            elif op == 0x87:
                b = st.pop() & 63; x = st[-1]
                x = x - (1 << 64) if x >> 63 else x
                st[-1] = (x >> b) & M64
            elif op == 0xAC:
                x = st[-1]
                st[-1] = (x | (M64 ^ M32)) if x & 0x80000000 else x
            elif op == 0x00:
                raise RuntimeError("wasm unreachable")
            elif op == 0x01:
                pass
            else:
                raise NotImplementedError("opcode 0x%02x" % op)
        self.steps += pc
        return st[-1] if nres else None


def fnv1a64(name: str):
    h = 0xCBF29CE484222325
    for ch in name.encode():
        h ^= ch
        h = (h * 0x100000001B3) & M64
    return h >> 32, h & M32


def calculate_witness(wasm_bytes: bytes, inputs: dict, sanity_check: int = 0):
    """inputs: {signal name: int or list of ints}.  Returns the witness as a list of canonical ints."""
    msgs = []

    def host(name, args):
        if name == "exceptionHandler":
            raise RuntimeError("circom exception %r %s" % (args, msgs))
        msgs.append((name, args))
        return None

    inst = Instance(Module(wasm_bytes), host)
    inst.call("init", sanity_check)
    n32 = inst.call("getFieldNumLen32")
    inst.call("getRawPrime")
    prime = 0
    for j in range(n32):
        prime |= inst.call("readSharedRWMemory", j) << (32 * j)
    for name, vals in inputs.items():
        msb, lsb = fnv1a64(name)
        vals = vals if isinstance(vals, (list, tuple)) else [vals]
        for i, v in enumerate(vals):
            v %= prime
            for j in range(n32):
                inst.call("writeSharedRWMemory", j, (v >> (32 * j)) & M32)
            inst.call("setInputSignal", msb, lsb, i)
    n = inst.call("getWitnessSize")
    out = []
    for i in range(n):
        inst.call("getWitness", i)
        v = 0
        for j in range(n32):
            v |= inst.call("readSharedRWMemory", j) << (32 * j)
        out.append(v)
    return out, prime
