"""Minimal JVM class-file reader + <clinit> interpreter.

Used by tools/make_golden.py to decode the reference's snapshot fixtures
(/root/reference/needle-compiler/src/test/resources/snapshots/*.class) into plain data: the
BYTE_CLASSES runs, the STATES_* tables, the accepting states and the accelerator constants the
reference's DFAClassBuilder emitted for 12 regexes.  No JVM is needed: the
static initialiser only pushes constants, allocates arrays and calls four known static helpers, which
are restated here (ByteClassUtil.fillBytes / fillMultipleByteClassesFromString*_singleArray,
needle-types/.../ByteClassUtil.java:44-120, and java.util.Arrays.fill).
"""
import struct


class ClassFile:
    def __init__(self, data: bytes):
        self.d = data
        self.pos = 0
        assert self.u4() == 0xCAFEBABE
        self.u2(), self.u2()  # minor, major
        n = self.u2()
        self.cp = [None] * n
        i = 1
        while i < n:
            tag = self.u1()
            if tag == 1:
                ln = self.u2()
                self.cp[i] = ("utf8", self._mutf8(self.d[self.pos:self.pos + ln]))
                self.pos += ln
            elif tag == 3:
                self.cp[i] = ("int", struct.unpack(">i", self.take(4))[0])
            elif tag == 4:
                self.cp[i] = ("float", struct.unpack(">f", self.take(4))[0])
            elif tag in (5, 6):
                self.cp[i] = ("wide", self.take(8))
                i += 1
            elif tag == 7:
                self.cp[i] = ("class", self.u2())
            elif tag == 8:
                self.cp[i] = ("string", self.u2())
            elif tag in (9, 10, 11):
                self.cp[i] = ({9: "field", 10: "method", 11: "imethod"}[tag], self.u2(), self.u2())
            elif tag == 12:
                self.cp[i] = ("nat", self.u2(), self.u2())
            elif tag == 15:
                self.cp[i] = ("mh", self.u1(), self.u2())
            elif tag == 16:
                self.cp[i] = ("mt", self.u2())
            elif tag == 18:
                self.cp[i] = ("indy", self.u2(), self.u2())
            else:
                raise ValueError(f"unknown constant pool tag {tag}")
            i += 1
        self.access = self.u2()
        self.this_class = self.u2()
        self.super_class = self.u2()
        self.interfaces = [self.u2() for _ in range(self.u2())]
        self.fields = [self._member() for _ in range(self.u2())]
        self.methods = [self._member() for _ in range(self.u2())]

    # -- primitives
    def take(self, n):
        b = self.d[self.pos:self.pos + n]
        self.pos += n
        return b

    def u1(self):
        return self.take(1)[0]

    def u2(self):
        return struct.unpack(">H", self.take(2))[0]

    def u4(self):
        return struct.unpack(">I", self.take(4))[0]

    @staticmethod
    def _mutf8(b: bytes) -> str:
        # modified UTF-8: BMP chars as 1-3 bytes, U+0000 as C0 80, no 4-byte forms
        out, i = [], 0
        while i < len(b):
            c = b[i]
            if c < 0x80:
                out.append(c)
                i += 1
            elif (c & 0xE0) == 0xC0:
                out.append(((c & 0x1F) << 6) | (b[i + 1] & 0x3F))
                i += 2
            else:
                out.append(((c & 0x0F) << 12) | ((b[i + 1] & 0x3F) << 6) | (b[i + 2] & 0x3F))
                i += 3
        return "".join(chr(x) for x in out)

    def utf8(self, idx):
        return self.cp[idx][1]

    def _member(self):
        access, name, desc = self.u2(), self.u2(), self.u2()
        attrs = {}
        for _ in range(self.u2()):
            an = self.utf8(self.u2())
            ln = self.u4()
            attrs[an] = self.take(ln)
        return {"access": access, "name": self.utf8(name), "desc": self.utf8(desc), "attrs": attrs}

    def const(self, idx):
        e = self.cp[idx]
        if e[0] == "int":
            return e[1]
        if e[0] == "string":
            return self.utf8(e[1])
        raise ValueError(f"unsupported ldc constant {e}")

    def ref(self, idx):
        kind, cls, nat = self.cp[idx]
        _, n, d = self.cp[nat]
        return self.utf8(self.cp[cls][1]), self.utf8(n), self.utf8(d)

    def field_constant(self, f):
        a = f["attrs"].get("ConstantValue")
        if a is None:
            return None
        return self.const(struct.unpack(">H", a)[0])

    def code(self, m):
        a = m["attrs"]["Code"]
        ln = struct.unpack(">I", a[4:8])[0]
        return a[8:8 + ln]


def _fill_multiple(table, stride, s):
    # ByteClassUtil.fillMultipleByteClassesFromString(UsingShorts)_singleArray (ByteClassUtil.java:50-120)
    for state_string in s.split(";"):
        state, transitions = state_string.split(":")
        state = int(state, 16)
        for comp in transitions.split(","):
            cls, target = comp.split("-")
            table[state * stride + int(cls, 16)] = int(target, 16)


def run_clinit(cf: ClassFile):
    """Interpret every <clinit> and return the static fields {name: value}."""
    statics = {}
    for f in cf.fields:
        v = cf.field_constant(f)
        if v is not None:
            statics[f["name"]] = v
    strides = {}
    for m in cf.methods:
        if m["name"] != "<clinit>":
            continue
        code = cf.code(m)
        pc, stack = 0, []
        while pc < len(code):
            op = code[pc]
            if 2 <= op <= 8:  # iconst_m1..iconst_5
                stack.append(op - 3)
                pc += 1
            elif op == 0x10:  # bipush
                stack.append(struct.unpack(">b", code[pc + 1:pc + 2])[0])
                pc += 2
            elif op == 0x11:  # sipush
                stack.append(struct.unpack(">h", code[pc + 1:pc + 3])[0])
                pc += 3
            elif op == 0x12:  # ldc
                stack.append(cf.const(code[pc + 1]))
                pc += 2
            elif op == 0x13:  # ldc_w
                stack.append(cf.const(struct.unpack(">H", code[pc + 1:pc + 3])[0]))
                pc += 3
            elif op == 0xBC:  # newarray
                n = stack.pop()
                stack.append({"type": code[pc + 1], "data": [0] * n})
                pc += 2
            elif op == 0x59:  # dup
                stack.append(stack[-1])
                pc += 1
            elif op in (0x54, 0x56, 0x4F):  # bastore, sastore, iastore
                v, i, arr = stack.pop(), stack.pop(), stack.pop()
                arr["data"][i] = v
                pc += 1
            elif op == 0xB3:  # putstatic
                _, name, _ = cf.ref(struct.unpack(">H", code[pc + 1:pc + 3])[0])
                statics[name] = stack.pop()
                pc += 3
            elif op == 0xB2:  # getstatic
                _, name, _ = cf.ref(struct.unpack(">H", code[pc + 1:pc + 3])[0])
                stack.append(statics[name])
                pc += 3
            elif op == 0xB8:  # invokestatic
                owner, name, desc = cf.ref(struct.unpack(">H", code[pc + 1:pc + 3])[0])
                if name == "fill":
                    v, arr = stack.pop(), stack.pop()
                    arr["data"][:] = [v] * len(arr["data"])
                elif name == "fillBytes":
                    hi, lo, v, arr = stack.pop(), stack.pop(), stack.pop(), stack.pop()
                    for i in range(lo, hi + 1):
                        arr["data"][i] = v
                    arr.setdefault("runs", []).append((v, lo, hi))
                elif name.startswith("fillMultipleByteClassesFromString"):
                    s, stride, arr = stack.pop(), stack.pop(), stack.pop()
                    _fill_multiple(arr["data"], stride, s)
                    arr["stride"] = stride
                    arr.setdefault("strings", []).append(s)
                else:
                    raise ValueError(f"unexpected static call {owner}.{name}{desc}")
                pc += 3
            elif op == 0xB1:  # return
                pc += 1
            else:
                raise ValueError(f"unhandled opcode 0x{op:02x} at pc={pc} in <clinit>")
    return statics


def int_constants(cf: ClassFile, m):
    """All int constants pushed by a method, in order (a light linear sweep good enough for the
    straight-line prologues of the generated methods)."""
    lengths = {0x10: 2, 0x11: 3, 0x12: 2, 0x13: 3, 0x14: 3, 0x15: 2, 0x16: 2, 0x17: 2, 0x18: 2, 0x19: 2, 0x36: 2, 0x37: 2,
               0x38: 2, 0x39: 2, 0x3A: 2, 0x84: 3, 0xBC: 2, 0xBD: 3, 0xC0: 3, 0xC1: 3, 0xBB: 3,
               0xB2: 3, 0xB3: 3, 0xB4: 3, 0xB5: 3, 0xB6: 3, 0xB7: 3, 0xB8: 3, 0xB9: 5, 0xBA: 5,
               0xC6: 3, 0xC7: 3, 0xC8: 5, 0xC9: 5, 0xA7: 3, 0xA8: 3, 0xA9: 2}
    for o in range(0x99, 0xA7):
        lengths[o] = 3  # if<cond>, if_icmp<cond>, if_acmp<cond>
    code = cf.code(m)
    pc, out = 0, []
    while pc < len(code):
        op = code[pc]
        if 2 <= op <= 8:
            out.append(op - 3)
        elif op == 0x10:
            out.append(struct.unpack(">b", code[pc + 1:pc + 2])[0])
        elif op == 0x11:
            out.append(struct.unpack(">h", code[pc + 1:pc + 3])[0])
        elif op in (0x12, 0x13):
            idx = code[pc + 1] if op == 0x12 else struct.unpack(">H", code[pc + 1:pc + 3])[0]
            if cf.cp[idx][0] == "int":
                out.append(cf.cp[idx][1])
        if op == 0xAA:  # tableswitch
            p = (pc + 4) & ~3
            lo, hi = struct.unpack(">ii", code[p + 4:p + 12])
            pc = p + 12 + 4 * (hi - lo + 1)
            continue
        if op == 0xAB:  # lookupswitch
            p = (pc + 4) & ~3
            n = struct.unpack(">i", code[p + 4:p + 8])[0]
            pc = p + 8 + 8 * n
            continue
        pc += lengths.get(op, 1)
    return out
