#!/usr/bin/env python
"""f2c_lite.py -- mechanical Fortran-90-subset -> C translator used to build oracle/_ref (TEST INFRASTRUCTURE ONLY).

There is no Fortran compiler in this image (gfortran / flang / nvfortran absent), so the reference cannot be compiled
as it is.  Its hot-path subroutines are, however, written in a very small subset of Fortran 90: scalar assignments,
do loops, if blocks, calls, module globals and arrays with explicit bounds.  This script translates exactly that subset,
statement by statement, from the reference's OWN source files where they lie under /root/reference into one C file
under oracle/_ref/ (git-ignored; never committed), which gcc then compiles with -ffp-contract=off.  Nothing of the
reference's arithmetic is restated by hand here: module variables, parameters, derived types, array bounds (the
reference's own allocate statements) and every expression come out of the reference's text.  Expression trees are emitted
fully parenthesised in Fortran's precedence / associativity, so evaluation order is the source order.

What it is for: pinning the hand-written restatement (oracle/mflbm_oracle.c) -- and through it the CUDA kernels -- to
the reference's real source (tests/test_ref_pin.py), and the `--impl reference` arm of bench.py.

Anything outside the subset raises Unsupported; the subroutine is then left out (calls to it become run-time aborts).
"""
import argparse
import os
import re
import subprocess
import sys


class Unsupported(Exception):
    pass


# ----------------------------------------------------------------------------------------------------------------
# source -> logical statements
# ----------------------------------------------------------------------------------------------------------------
def preprocess(path, defines=()):
    """cpp in traditional mode, like gfortran does for .F90 files"""
    cmd = ["gcc", "-E", "-P", "-traditional-cpp", "-x", "c", "-I", os.path.dirname(path)] + ["-D" + d for d in defines] + [path]
    return subprocess.run(cmd, check=True, capture_output=True, text=True).stdout


def strip_comment(line):
    out, q = [], None
    for ch in line:
        if q:
            out.append(ch)
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            out.append(ch)
        elif ch == "!":
            break
        else:
            out.append(ch)
    return "".join(out)


def logical_lines(text):
    """join continuation lines, drop comments, keep !$omp directives as ('omp', text) entries; lower-case outside strings"""
    res, cur, omp = [], "", ""
    for raw in text.split("\n"):
        s = raw.strip()
        if not s:
            continue
        low = s.lower()
        if low.startswith("!$omp"):
            body = s[5:].strip()
            if body.startswith("&"):
                body = body[1:].strip()
            cont = body.endswith("&")
            if cont:
                body = body[:-1]
            omp += " " + body
            if not cont:
                res.append(("omp", lower_outside_strings(omp.strip())))
                omp = ""
            continue
        if s.startswith("!"):
            continue
        s = strip_comment(s).strip()
        if not s:
            continue
        if s.startswith("&"):
            s = s[1:]
        if s.endswith("&"):
            cur += s[:-1] + " "
            continue
        cur += s
        for part in split_semicolons(cur):
            if part.strip():
                res.append(("stmt", lower_outside_strings(part.strip())))
        cur = ""
    return res


def split_semicolons(s):
    out, cur, q = [], "", None
    for ch in s:
        if q:
            cur += ch
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            cur += ch
        elif ch == ";":
            out.append(cur)
            cur = ""
        else:
            cur += ch
    out.append(cur)
    return out


def lower_outside_strings(s):
    out, q = [], None
    for ch in s:
        if q:
            out.append(ch)
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            out.append(ch)
        else:
            out.append(ch.lower())
    return "".join(out)


# ----------------------------------------------------------------------------------------------------------------
# expressions
# ----------------------------------------------------------------------------------------------------------------
TOKEN = re.compile(r"""\s*(?:
    (?P<num>(?:\d+\.(?![a-z]+\.)\d*|\.\d+|\d+)(?:[ed][+-]?\d+)?(?:_\w+)?) |
    (?P<dotop>\.(?:and|or|not|eq|ne|lt|le|gt|ge|eqv|neqv|true|false)\.) |
    (?P<id>[a-z_]\w*) |
    (?P<str>'(?:[^']|'')*'|"(?:[^"]|"")*") |
    (?P<op>\*\*|==|/=|<=|>=|=>|\(/|/\)|//|[-+*/(),<>=%:])
)""", re.X)


def tokenize(s):
    toks, pos = [], 0
    s = s.rstrip()
    while pos < len(s):
        m = TOKEN.match(s, pos)
        if not m or m.end() == pos:
            raise Unsupported("cannot tokenize: %r" % s[pos:pos + 30])
        pos = m.end()
        kind = m.lastgroup
        toks.append((kind, m.group(kind)))
    return toks


INTRINSIC = {"dsqrt": "sqrt", "sqrt": "sqrt", "dabs": "f_abs", "abs": "f_abs", "iabs": "f_abs", "dcos": "cos", "cos": "cos", "dsin": "sin",
             "sin": "sin", "dtan": "tan", "tan": "tan", "dexp": "exp", "exp": "exp", "dlog": "log", "log": "log", "dtanh": "tanh",
             "tanh": "tanh", "dacos": "acos", "acos": "acos", "datan": "atan", "atan": "atan", "dble": "f_dble", "mod": "f_mod",
             "max": "f_max", "min": "f_min", "dmax1": "f_max", "dmin1": "f_min", "int": "f_int", "nint": "f_nint", "floor": "f_floor",
             "ceiling": "f_ceiling", "real": "f_real", "isnan": "f_isnan", "sign": "f_sign", "dsign": "f_sign", "idnint": "f_nint"}


class Parser:
    """Fortran expression -> fully parenthesised C text"""

    def __init__(self, toks, ctx):
        self.t, self.i, self.ctx = toks, 0, ctx

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else (None, None)

    def next(self):
        tok = self.peek()
        self.i += 1
        return tok

    def expect(self, v):
        k, val = self.next()
        if val != v:
            raise Unsupported("expected %r, got %r" % (v, val))

    def parse(self):
        e = self.equiv()
        if self.i != len(self.t):
            raise Unsupported("trailing tokens %r" % (self.t[self.i:],))
        return e

    def equiv(self):
        l = self.orx()
        while self.peek()[1] in (".eqv.", ".neqv."):
            op = self.next()[1]
            r = self.orx()
            l = "((!!(%s)) %s (!!(%s)))" % (l, "==" if op == ".eqv." else "!=", r)
        return l

    def orx(self):
        l = self.andx()
        while self.peek()[1] == ".or.":
            self.next()
            l = "(%s || %s)" % (l, self.andx())
        return l

    def andx(self):
        l = self.notx()
        while self.peek()[1] == ".and.":
            self.next()
            l = "(%s && %s)" % (l, self.notx())
        return l

    def notx(self):
        if self.peek()[1] == ".not.":
            self.next()
            return "(!%s)" % self.notx()
        return self.rel()

    REL = {"==": "==", "/=": "!=", "<": "<", "<=": "<=", ">": ">", ">=": ">=", ".eq.": "==", ".ne.": "!=", ".lt.": "<", ".le.": "<=",
           ".gt.": ">", ".ge.": ">="}

    def rel(self):
        l = self.add()
        if self.peek()[1] in self.REL:
            op = self.REL[self.next()[1]]
            l = "(%s %s %s)" % (l, op, self.add())
        return l

    def add(self):
        k, v = self.peek()
        if v in ("+", "-"):  # leading sign applies to the first TERM (Fortran: -a*b = -(a*b))
            self.next()
            l = self.mul()
            l = "(-%s)" % l if v == "-" else l
        else:
            l = self.mul()
        while self.peek()[1] in ("+", "-"):
            op = self.next()[1]
            l = "(%s %s %s)" % (l, op, self.mul())
        return l

    def mul(self):
        l = self.power()
        while self.peek()[1] in ("*", "/"):
            op = self.next()[1]
            l = "(%s %s %s)" % (l, op, self.power())
        return l

    def power(self):
        base = self.primary()
        if self.peek()[1] == "**":
            self.next()
            k, v = self.peek()
            if v in ("+", "-"):
                self.next()
                e = self.power()
                e = "(-%s)" % e if v == "-" else e
            else:
                e = self.power()  # right associative
            m = re.fullmatch(r"\(?(\d+)\)?", e)
            if m and re.fullmatch(r"[a-z_]\w*", base) and self.ctx.is_int(base):
                return "f_ipow(%s, %d)" % (base, int(m.group(1)))  # integer ** integer stays integer arithmetic (and wraps)
            if m:  # integer power: repeated multiplication like gfortran's expansion of x**2 / x**3
                n = int(m.group(1))
                if n == 2:
                    return "f_sq(%s)" % base
                if n == 3:
                    return "f_cube(%s)" % base
                return "f_powi(%s, %d)" % (base, n)
            return "pow(%s, %s)" % (base, e)
        return base

    def primary(self):
        k, v = self.next()
        if k == "num":
            return self.number(v)
        if k == "str":
            return '"%s"' % v[1:-1].replace("\\", "\\\\").replace('"', '\\"')
        if k == "dotop":
            if v == ".true.":
                return "1"
            if v == ".false.":
                return "0"
            raise Unsupported("operator %s in primary position" % v)
        if v == "(":
            e = self.equiv()
            self.expect(")")
            return "(%s)" % e
        if v == "(/":
            raise Unsupported("array constructor in an expression")
        if k == "id":
            return self.designator(v)
        raise Unsupported("unexpected token %r" % v)

    def number(self, v):
        v = re.sub(r"_\w+$", "", v)
        if re.fullmatch(r"\d+", v):
            return v
        if "d" in v:
            return v.replace("d", "e")
        return v + "f"  # default-kind real literal: single precision, promoted like Fortran does

    def args(self):
        a = []
        if self.peek()[1] == ")":
            self.next()
            return a
        while True:
            if self.peek()[1] == ":":
                raise Unsupported("array section")
            a.append(self.equiv())
            k, v = self.next()
            if v == ")":
                return a
            if v == ":":
                raise Unsupported("array section")
            if v != ",":
                raise Unsupported("bad argument list near %r" % v)

    def designator(self, name):
        out = None
        while True:
            if self.peek()[1] == "(":
                self.next()
                a = self.args()
                if out is None and name in INTRINSIC and not self.ctx.is_array(name):
                    f = INTRINSIC[name]
                    if f in ("f_max", "f_min"):
                        if len(a) < 2:
                            raise Unsupported("max/min arity")
                        e = a[0]
                        for x in a[1:]:
                            e = "%s2(%s, %s)" % (f, e, x)
                        cur = e
                    elif f == "f_real" and len(a) == 2:
                        cur = "f_dble(%s)" % a[0]  # real(x, kind=8)
                    elif f == "f_int" and len(a) == 2:
                        cur = "f_int(%s)" % a[0]
                    else:
                        cur = "%s(%s)" % (f, ", ".join(a))
                else:
                    self.ctx.note_ref(name if out is None else None, len(a))
                    cur = "%s(%s)" % (self.ctx.cname(name) if out is None else name, ", ".join(a))
            else:
                if out is None:
                    self.ctx.note_scalar(name)
                cur = self.ctx.cname(name) if out is None else name
            out = cur if out is None else out + "." + cur
            if self.peek()[1] == "%":
                self.next()
                k, name = self.next()
                if k != "id":
                    raise Unsupported("bad component reference")
                continue
            return out


C_RESERVED = {"auto", "break", "case", "char", "const", "continue", "default", "do", "double", "else", "enum", "extern", "float", "for",
              "goto", "if", "inline", "int", "long", "register", "restrict", "return", "short", "signed", "sizeof", "static", "struct",
              "switch", "typedef", "union", "unsigned", "void", "volatile", "while", "main", "abs", "exp", "log", "sin", "cos", "tan", "pow",
              "sqrt", "free", "calloc", "printf", "abort", "index", "time"}


class Ctx:
    def __init__(self, arrays, ints=()):
        self.arrays = arrays  # names known to be arrays (module level + current locals)
        self.local_arrays = set()
        self.ints = set(ints)  # integer scalars (module level + current locals / arguments)

    def is_int(self, name):
        return name in self.ints

    def is_array(self, name):
        return name in self.arrays or name in self.local_arrays

    def cname(self, name):
        return name + "_v" if name in C_RESERVED else name

    def note_ref(self, name, nargs):
        if name is not None and not self.is_array(name):
            raise Unsupported("reference to unknown array / function %s(...)" % name)

    def note_scalar(self, name):
        pass


def cexpr(s, ctx):
    return Parser(tokenize(s), ctx).parse()


# ----------------------------------------------------------------------------------------------------------------
# declarations
# ----------------------------------------------------------------------------------------------------------------
def split_top(s, sep=","):
    out, depth, cur, q = [], 0, "", None
    i = 0
    while i < len(s):
        ch = s[i]
        if q:
            cur += ch
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            cur += ch
        elif s.startswith("(/", i):
            depth += 1
            cur += "(/"
            i += 1
        elif s.startswith("/)", i):
            depth -= 1
            cur += "/)"
            i += 1
        elif ch == "(":
            depth += 1
            cur += ch
        elif ch == ")":
            depth -= 1
            cur += ch
        elif ch == sep and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
        i += 1
    if cur.strip():
        out.append(cur.strip())
    return out


DECL = re.compile(r"^(integer|real|double\s+precision|logical|character|type\s*\(\s*\w+\s*\))\s*(\([^)]*\))?\s*(.*)$")


def ctype_of(base, kind):
    base = re.sub(r"\s+", " ", base)
    k = None
    if kind:
        m = re.search(r"(\d+)", kind)
        k = int(m.group(1)) if m else None
    if base == "integer":
        return {1: "signed char", 2: "short", 8: "long long"}.get(k, "int")
    if base == "real":
        return "float" if k in (None, 4) and k != 8 else "double" if k == 8 else "float"
    if base == "double precision":
        return "double"
    if base == "logical":
        return "int"
    if base.startswith("type"):
        return "struct " + re.search(r"\(\s*(\w+)\s*\)", base).group(1)
    return None  # character


def parse_decl(stmt):
    """-> None or dict(ctype, attrs{parameter, allocatable, dimension}, entities [(name, dims|None, init|None)])"""
    m = DECL.match(stmt)
    if not m:
        return None
    base, kind, rest = m.group(1), m.group(2), m.group(3)
    if base.startswith("type") and "::" not in rest and not rest.startswith(","):
        return None
    if base == "character":
        return dict(ctype=None, attrs={}, entities=[])
    # "real(kind=8)" puts the kind into group 2; "integer, parameter :: x" has none
    attrs = {}
    if "::" in rest:
        a, ents = rest.split("::", 1)
        for at in split_top(a.strip().lstrip(",")):
            at = at.strip()
            if not at:
                continue
            mm = re.match(r"dimension\s*\((.*)\)$", at)
            if mm:
                attrs["dimension"] = mm.group(1)
            else:
                attrs[at.split("(")[0].strip()] = at
    else:
        ents = rest
    entities = []
    for e in split_top(ents):
        init = None
        if "=" in e and not re.match(r"^[^=]*\([^)]*=", e):
            # split at the first top-level '='
            depth = 0
            for i, ch in enumerate(e):
                if ch == "(":
                    depth += 1
                elif ch == ")":
                    depth -= 1
                elif ch == "=" and depth == 0:
                    init = e[i + 1:].strip()
                    e = e[:i].strip()
                    break
        mm = re.match(r"^(\w+)\s*(?:\((.*)\))?$", e.strip())
        if not mm:
            raise Unsupported("cannot parse declared entity %r" % e)
        entities.append((mm.group(1), mm.group(2) or attrs.get("dimension"), init))
    return dict(ctype=ctype_of(base, kind), attrs=attrs, entities=entities)


# ----------------------------------------------------------------------------------------------------------------
# program units
# ----------------------------------------------------------------------------------------------------------------
class Module:
    def __init__(self):
        self.scalars = {}   # name -> ctype
        self.params = []    # (name, ctype, init C text)
        self.parrays = []   # (name, ctype, lo, [values])
        self.arrays = {}    # name -> (ctype, rank)
        self.types = {}     # type name -> [(member, ctype, dims)]
        self.order = []


def parse_modules(lines, mod):
    """module-level declarations of every module in `lines`"""
    ctx = Ctx(set())
    in_mod, in_type = False, None
    for kind, s in lines:
        if kind != "stmt":
            continue
        if re.match(r"^module\s+\w+$", s) and not s.startswith("module procedure"):
            in_mod = True
            continue
        if re.match(r"^end\s*module", s):
            in_mod = False
            continue
        if not in_mod:
            continue
        m = re.match(r"^type\s+(\w+)$", s)
        if m:
            in_type = m.group(1)
            mod.types[in_type] = []
            continue
        if re.match(r"^end\s*type", s):
            in_type = None
            continue
        if s in ("implicit none", "save") or s.startswith("use ") or s.startswith("character"):
            continue
        d = parse_decl(s)
        if d is None:
            raise Unsupported("module-level statement %r" % s)
        if d["ctype"] is None:
            continue
        for name, dims, init in d["entities"]:
            if in_type:
                mod.types[in_type].append((name, d["ctype"], dims))
                ctx.arrays.add(name) if dims else None
                continue
            if "parameter" in d["attrs"]:
                if dims:  # name(lo:hi) = (/ ... /)
                    lo = dims.split(":")[0] if ":" in dims else "1"
                    if dims.strip() == ":":
                        raise Unsupported("assumed-shape parameter array %s" % name)
                    mm = re.match(r"^\(/(.*)/\)$", init.strip(), re.S)
                    vals = [cexpr(v, ctx) for v in split_top(mm.group(1))]
                    mod.parrays.append((name, d["ctype"], cexpr(lo, ctx), vals))
                    ctx.arrays.add(name)
                else:
                    mod.params.append((name, d["ctype"], cexpr(init, ctx)))
            elif dims:
                rank = len(split_top(dims))
                mod.arrays[name] = (d["ctype"], rank, None if "allocatable" in d["attrs"] or ":" == dims.strip()[0] else dims)
                ctx.arrays.add(name)
            else:
                mod.scalars[name] = d["ctype"]
    # "dimension(:), parameter :: ex(0:18) = ..." puts the real bounds on the entity: handled above through `dims`
    return mod


class Sub:
    def __init__(self, name, args):
        self.name, self.args = name, args
        self.body = []  # (kind, stmt)


def split_subroutines(lines):
    subs, cur = [], None
    for kind, s in lines:
        if kind == "stmt":
            m = re.match(r"^subroutine\s+(\w+)\s*(?:\((.*)\))?$", s)
            if m:
                cur = Sub(m.group(1), [a.strip() for a in m.group(2).split(",")] if m.group(2) and m.group(2).strip() else [])
                subs.append(cur)
                continue
            if cur is not None and (re.match(r"^end\s*subroutine", s) or s == "end"):
                cur = None
                continue
        if cur is not None:
            cur.body.append((kind, s))
    return subs


def omp_pragma(text, ctx):
    """'parallel do ... private(list) collapse(n) reduction(op:list)' -> '#pragma omp parallel for ...' (None if not a loop directive)"""
    t = text.strip()
    if not t.startswith("parallel do"):
        return None
    out = ["#pragma omp parallel for"]
    for m in re.finditer(r"(private|firstprivate|reduction|collapse|schedule)\s*\(([^)]*)\)", t):
        cl, arg = m.group(1), m.group(2)
        if cl in ("private", "firstprivate"):
            names = [ctx.cname(x.strip()) for x in arg.replace("&", " ").split(",") if x.strip()]
            out.append("%s(%s)" % (cl, ", ".join(names)))
        elif cl == "reduction":
            op, names = arg.split(":")
            op = {"+": "+", "*": "*", "max": "max", "min": "min", ".or.": "||", ".and.": "&&"}[op.strip()]
            out.append("reduction(%s: %s)" % (op, ", ".join(ctx.cname(x.strip()) for x in names.split(","))))
        else:
            out.append("%s(%s)" % (cl, arg))
    return " ".join(out)


def translate_sub(sub, mod, known_subs, openmp):
    ctx = Ctx(set(mod.arrays) | {n for n, *_ in mod.parrays} | {m for t in mod.types.values() for m, _, d in t if d},
              {n for n, ct in mod.scalars.items() if ct in ("int", "signed char", "short", "long long")})
    decls, code, ind = [], [], 1
    argtypes = {}
    locals_ = {}
    assigned = set()
    pending_omp = None
    stack = []

    def emit(s):
        code.append("    " * ind + s)

    def one(kind, s):
        nonlocal ind, pending_omp
        if kind == "omp":
            if openmp:
                p = omp_pragma(s, ctx)
                if p:
                    pending_omp = p
            return
        if s.startswith("use ") or s.startswith("use,") or s == "implicit none" or s.startswith("include ") or s == "save" or s.startswith("external "):
            return
        m = re.match(r"^(integer|logical|real\s*\(kind=8\)|double precision)\s+([a-z_]\w*(?:\s*,\s*[a-z_]\w*)*)$", s)
        if m and "::" not in s:  # old-style declaration without '::'
            s = "%s :: %s" % (m.group(1), m.group(2))
        if s.startswith("character"):
            return
        d = parse_decl(s) if "::" in s or DECL.match(s) and not re.match(r"^(integer|real|logical)\s*\(", s.split("=")[0] if "=" in s and "::" not in s else "x") else None
        if d is not None and ("::" in s):
            if d["ctype"] is None:
                return
            for name, dims, init in d["entities"]:
                if not dims and d["ctype"] in ("int", "signed char", "short", "long long"):
                    ctx.ints.add(name)
                else:
                    ctx.ints.discard(name)  # a local of another type shadows a module integer
                if name in sub.args:
                    if dims:
                        raise Unsupported("array dummy argument %s" % name)
                    argtypes[name] = d["ctype"]
                    continue
                if "parameter" in d["attrs"]:
                    if dims:
                        raise Unsupported("local parameter array")
                    decls.append("const %s %s = %s;" % (d["ctype"], ctx.cname(name), cexpr(init, ctx)))
                    continue
                if dims:
                    bounds = []
                    for b in split_top(dims):
                        lo, hi = (b.split(":") + [None])[:2] if ":" in b else ("1", b)
                        bounds.append((cexpr(lo, ctx), cexpr(hi, ctx)))
                    ctx.local_arrays.add(name)
                    locals_[name] = (d["ctype"], bounds)
                else:
                    locals_[name] = (d["ctype"], None)
                    if init is not None:
                        raise Unsupported("initialised local %s (implies SAVE)" % name)
            return
        # ---- executable statements ----
        if s == "return":
            emit("return;")
            return
        if s in ("continue",):
            return
        if s == "exit":
            emit("break;")
            return
        if s == "cycle":
            emit("continue;")
            return
        if s.startswith("stop"):
            emit("ref_abort(\"%s: stop\");" % sub.name)
            return
        m = re.match(r"^do\s+(\w+)\s*=\s*(.*)$", s)
        if m:
            v = ctx.cname(m.group(1))
            parts = split_top(m.group(2))
            lo, hi = cexpr(parts[0], ctx), cexpr(parts[1], ctx)
            if pending_omp:
                code.append(pending_omp)
                pending_omp = None
            if len(parts) == 3:
                st = cexpr(parts[2], ctx)
                emit("for (%s = %s; (%s) > 0 ? %s <= %s : %s >= %s; %s += %s) {" % (v, lo, st, v, hi, v, hi, v, st))
            else:
                emit("for (%s = %s; %s <= %s; %s++) {" % (v, lo, v, hi, v))
            assigned.add(m.group(1))
            ind += 1
            stack.append("do")
            return
        m = re.match(r"^do\s+while\s*\((.*)\)$", s)
        if m:
            emit("while (%s) {" % cexpr(m.group(1), ctx))
            ind += 1
            stack.append("do")
            return
        if re.match(r"^end\s*do$", s):
            ind -= 1
            emit("}")
            stack.pop()
            return
        m = re.match(r"^if\s*\((.*)\)\s*then$", s)
        if m:
            emit("if (%s) {" % cexpr(m.group(1), ctx))
            ind += 1
            stack.append("if")
            return
        m = re.match(r"^else\s*if\s*\((.*)\)\s*then$", s)
        if m:
            ind -= 1
            emit("} else if (%s) {" % cexpr(m.group(1), ctx))
            ind += 1
            return
        if s == "else":
            ind -= 1
            emit("} else {")
            ind += 1
            return
        if re.match(r"^end\s*if$", s):
            ind -= 1
            emit("}")
            stack.pop()
            return
        pending_omp = None
        m = re.match(r"^if\s*\(", s)
        if m:  # one-line if: find the matching parenthesis
            depth, j = 0, s.index("(")
            for j in range(s.index("("), len(s)):
                depth += s[j] == "("
                depth -= s[j] == ")"
                if depth == 0:
                    break
            cond, rest = s[s.index("(") + 1:j], s[j + 1:].strip()
            inner = translate_simple(rest, ctx, sub, known_subs, assigned)
            emit("if (%s) { %s }" % (cexpr(cond, ctx), inner))
            return
        try:
            emit(translate_simple(s, ctx, sub, known_subs, assigned))
        except Unsupported as e:
            raise Unsupported("%s  [in: %s]" % (e, s[:80]))

    for kind, s in sub.body:
        try:
            one(kind, s)
        except Unsupported as e:
            raise Unsupported(str(e) if "[in:" in str(e) else "%s  [in: %s]" % (e, s[:90]))
    if stack:
        raise Unsupported("unbalanced blocks in %s" % sub.name)
    for a in sub.args:
        if a in assigned:
            raise Unsupported("%s assigns its dummy argument %s (arguments are passed by value here)" % (sub.name, a))
        if a not in argtypes:
            raise Unsupported("%s: dummy argument %s has no declaration" % (sub.name, a))
    head = "static void %s(%s)" % (sub.name, ", ".join("%s %s" % (argtypes[a], ctx.cname(a)) for a in sub.args) or "void")
    for name, (ct, bounds) in locals_.items():
        if bounds is None:
            decls.append("%s %s = 0;" % (ct, ctx.cname(name)))
        else:
            raise Unsupported("local array %s" % name)
    body = "\n".join("    " + d for d in decls) + "\n" + "\n".join(code)
    return head, body, [argtypes[a] for a in sub.args]


def translate_simple(s, ctx, sub, known_subs, assigned):
    m = re.match(r"^call\s+(\w+)\s*(?:\((.*)\))?$", s)
    if m:
        name, args = m.group(1), m.group(2)
        a = [cexpr(x, ctx) for x in split_top(args)] if args and args.strip() else []
        if name == "random_seed":
            return "/* random_seed(): unseeded in the reference; random_number itself is not provided */;"
        if name in ("mpi_barrier", "mpi_bcast"):
            return "/* %s: nothing to do in a single process */;" % name
        if name in ("mpi_reduce", "mpi_allreduce"):  # one rank: the reduction of a scalar is a copy (arrays: not supported)
            if len(a) >= 3 and a[2] == "1" and not ctx.is_array(split_top(args)[1].strip()):
                return "%s = %s; /* %s over one rank */" % (a[1], a[0], name)
            raise Unsupported("%s of an array" % name)
        known_subs.setdefault("__called__", set()).add(name)
        return "%s(%s);" % (name, ", ".join(a))
    if re.match(r"^(print|write)\b", s):
        return "/* %s */;" % re.sub(r"\*/", "* /", s[:60])
    m = re.match(r"^allocate\s*\((.*)\)$", s)
    if m:
        out = []
        for ent in split_top(m.group(1)):
            mm = re.match(r"^(\w+)\s*\((.*)\)$", ent)
            if not mm:
                raise Unsupported("allocate %r" % ent)
            b = []
            for dim in split_top(mm.group(2)):
                lo, hi = dim.split(":") if ":" in dim else ("1", dim)
                b += [cexpr(lo, ctx), cexpr(hi, ctx)]
            out.append("REF_ALLOC%d(%s, %s);" % (len(b) // 2, mm.group(1), ", ".join(b)))
        return " ".join(out)
    m = re.match(r"^deallocate\s*\((.*)\)$", s)
    if m:
        return " ".join("REF_FREE(%s);" % x.strip() for x in split_top(m.group(1)))
    if re.match(r"^(open|close|read|rewind|format|goto|select|where|forall|data|common|entry|interface)\b", s):
        raise Unsupported("statement %r" % s[:40])
    # assignment: split at the top-level '=' that is not part of ==, /=, <=, >=
    depth = 0
    for i, ch in enumerate(s):
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        elif ch == "=" and depth == 0 and s[i + 1:i + 2] != "=" and s[i - 1] not in "=/<>":
            lhs, rhs = s[:i].strip(), s[i + 1:].strip()
            mm = re.match(r"^(\w+)", lhs)
            base = mm.group(1)
            if re.fullmatch(r"\w+", lhs):
                if ctx.is_array(lhs):
                    raise Unsupported("whole-array assignment to %s" % lhs)
                assigned.add(lhs)
            return "%s = %s;" % (cexpr(lhs, ctx), cexpr(rhs, ctx))
    raise Unsupported("statement %r" % s[:60])


# ----------------------------------------------------------------------------------------------------------------
# C emission
# ----------------------------------------------------------------------------------------------------------------
PRELUDE = r"""/* GENERATED by oracle/f2c_lite.py from the reference's Fortran sources -- do not edit, do not commit. */
typedef unsigned long size_t;
double sqrt(double); double cos(double); double sin(double); double tan(double); double exp(double); double log(double);
double tanh(double); double acos(double); double atan(double); double pow(double, double); double fabs(double); double fmod(double, double);
double floor(double); double ceil(double); double copysign(double, double);
void *calloc(size_t, size_t); void free(void *); int printf(const char *, ...); void abort(void); int strcmp(const char *, const char *);
static void ref_abort(const char *why) { printf("oracle/_ref: %s\n", why); abort(); }
#define mpi_status_size 6 /* mpif.h stand-in: only sizes the (unused) status arrays of MemAllocate_multi */
#define f_sq(x) ((x) * (x))
#define f_cube(x) ((x) * (x) * (x))
static int f_ipow(int b, int e) { unsigned r = 1u, x = (unsigned)b; while (e > 0) { if (e & 1) r *= x; x *= x; e >>= 1; } return (int)r; }
static double f_powi(double x, int n) { double r = 1.0; int m = n < 0 ? -n : n; while (m) { if (m & 1) r *= x; x *= x; m >>= 1; } return n < 0 ? 1.0 / r : r; }
#define f_abs(x) _Generic((x), int: f_iabs, signed char: f_iabs, short: f_iabs, long long: f_labs, float: f_fabsf, default: fabs)(x)
static int f_iabs(int x) { return x < 0 ? -x : x; }
static long long f_labs(long long x) { return x < 0 ? -x : x; }
static float f_fabsf(float x) { return x < 0 ? -x : x; }
#define f_mod(a, b) _Generic((a) + (b), int: f_imod, long long: f_lmod, default: fmod)(a, b)
static int f_imod(int a, int b) { return a % b; }
static long long f_lmod(long long a, long long b) { return a % b; }
#define f_max2(a, b) ((a) > (b) ? (a) : (b))
#define f_min2(a, b) ((a) < (b) ? (a) : (b))
#define f_dble(x) ((double)(x))
#define f_real(x) ((float)(x))
#define f_int(x) ((int)(x))
#define f_nint(x) ((int)((x) >= 0 ? floor((double)(x) + 0.5) : -floor(0.5 - (double)(x))))
#define f_floor(x) ((int)floor((double)(x)))
#define f_ceiling(x) ((int)ceil((double)(x)))
#define f_isnan(x) ((x) != (x))
#define f_sign(a, b) _Generic((a), int: f_isign, default: f_dsign)(a, b)
static int f_isign(int a, int b) { int m = a < 0 ? -a : a; return b >= 0 ? m : -m; }
static double f_dsign(double a, double b) { return copysign(fabs(a), b); }
typedef struct { const char *name; void *base; long long lo[4], n[4]; int rank, elem; } ref_array_desc;
#define REF_IDX1(a, i) ((size_t)((i) - a##_d.lo[0]))
#define REF_IDX2(a, i, j) ((size_t)((i) - a##_d.lo[0]) + (size_t)a##_d.n[0] * (size_t)((j) - a##_d.lo[1]))
#define REF_IDX3(a, i, j, k) ((size_t)((i) - a##_d.lo[0]) + (size_t)a##_d.n[0] * ((size_t)((j) - a##_d.lo[1]) + (size_t)a##_d.n[1] * (size_t)((k) - a##_d.lo[2])))
#define REF_SET(a, r, e) do { a##_d.rank = r; a##_d.elem = (int)(e); } while (0)
#define REF_DIM(a, m, l, h) do { a##_d.lo[m] = (l); a##_d.n[m] = (long long)(h) - (long long)(l) + 1; if (a##_d.n[m] < 0) a##_d.n[m] = 0; } while (0)
#define REF_DO_ALLOC(a, cnt) do { if (a##_) free(a##_); a##_ = calloc((size_t)(cnt) + 1, sizeof *a##_); a##_d.base = a##_; } while (0)
#define REF_ALLOC1(a, l0, h0) do { REF_DIM(a, 0, l0, h0); REF_SET(a, 1, sizeof *a##_); REF_DO_ALLOC(a, a##_d.n[0]); } while (0)
#define REF_ALLOC2(a, l0, h0, l1, h1) do { REF_DIM(a, 0, l0, h0); REF_DIM(a, 1, l1, h1); REF_SET(a, 2, sizeof *a##_); REF_DO_ALLOC(a, a##_d.n[0] * a##_d.n[1]); } while (0)
#define REF_ALLOC3(a, l0, h0, l1, h1, l2, h2) do { REF_DIM(a, 0, l0, h0); REF_DIM(a, 1, l1, h1); REF_DIM(a, 2, l2, h2); REF_SET(a, 3, sizeof *a##_); REF_DO_ALLOC(a, a##_d.n[0] * a##_d.n[1] * a##_d.n[2]); } while (0)
#define REF_FREE(a) do { free(a##_); a##_ = 0; a##_d.base = 0; } while (0)
"""

API = r"""
/* ---- exported registry API (oracle/ref.py binds these) ---- */
#define REF_EXPORT __attribute__((visibility("default")))
REF_EXPORT int ref_set_int(const char *name, long long v) {
    for (int i = 0; ref_scalars[i].name; i++) if (!strcmp(ref_scalars[i].name, name)) {
        switch (ref_scalars[i].kind) {
        case 0: *(int *)ref_scalars[i].p = (int)v; return 0;
        case 1: *(double *)ref_scalars[i].p = (double)v; return 0;
        case 2: *(signed char *)ref_scalars[i].p = (signed char)v; return 0;
        case 3: *(long long *)ref_scalars[i].p = v; return 0;
        case 4: *(float *)ref_scalars[i].p = (float)v; return 0;
        }
    }
    return -1;
}
REF_EXPORT int ref_set_double(const char *name, double v) {
    for (int i = 0; ref_scalars[i].name; i++) if (!strcmp(ref_scalars[i].name, name)) {
        if (ref_scalars[i].kind == 1) { *(double *)ref_scalars[i].p = v; return 0; }
        if (ref_scalars[i].kind == 4) { *(float *)ref_scalars[i].p = (float)v; return 0; }
        return -2;
    }
    return -1;
}
REF_EXPORT double ref_get(const char *name) {
    for (int i = 0; ref_scalars[i].name; i++) if (!strcmp(ref_scalars[i].name, name)) {
        switch (ref_scalars[i].kind) {
        case 0: return *(int *)ref_scalars[i].p;
        case 1: return *(double *)ref_scalars[i].p;
        case 2: return *(signed char *)ref_scalars[i].p;
        case 3: return (double)*(long long *)ref_scalars[i].p;
        case 4: return *(float *)ref_scalars[i].p;
        }
    }
    return 0.0 / 0.0;
}
REF_EXPORT int ref_has(const char *name) {
    for (int i = 0; ref_scalars[i].name; i++) if (!strcmp(ref_scalars[i].name, name)) return 1;
    for (int i = 0; ref_arrays[i]; i++) if (!strcmp(ref_arrays[i]->name, name)) return 2;
    for (int i = 0; ref_subs[i].name; i++) if (!strcmp(ref_subs[i].name, name)) return 3;
    return 0;
}
REF_EXPORT const ref_array_desc *ref_array(const char *name) {
    for (int i = 0; ref_arrays[i]; i++) if (!strcmp(ref_arrays[i]->name, name)) return ref_arrays[i];
    return 0;
}
REF_EXPORT int ref_alloc(const char *name, int rank, const long long *lo, const long long *hi) {
    for (int i = 0; ref_arrays[i]; i++) if (!strcmp(ref_arrays[i]->name, name)) { ref_allocators[i](rank, lo, hi); return 0; }
    return -1;
}
REF_EXPORT int ref_call(const char *name, const long long *ia, const double *da) {
    for (int i = 0; ref_subs[i].name; i++) if (!strcmp(ref_subs[i].name, name)) { ref_subs[i].fn(ia, da); return 0; }
    return -1;
}
REF_EXPORT const char *ref_sub_name(int i) { return ref_subs[i].name; }
"""

KIND = {"int": 0, "double": 1, "signed char": 2, "long long": 3, "float": 4, "short": 0}


def emit_c(mod, subs_c, skipped, called, sources):
    o = [PRELUDE]
    o.append("/* sources: %s */" % ", ".join(sources))
    for tname, members in mod.types.items():
        o.append("struct %s {" % tname)
        for name, ct, dims in members:
            if dims:
                b = split_top(dims)
                if len(b) != 1:
                    raise Unsupported("multi-dimensional type member")
                lo, hi = b[0].split(":") if ":" in b[0] else ("1", b[0])
                o.append("    %s %s_[(%s) - (%s) + 1];" % (ct, name, hi, lo))
                o.append("#define %s(i) %s_[(i) - (%s)]" % (name, name, lo))
            else:
                o.append("    %s %s;" % (ct, name))
        o.append("};")
    ctxn = Ctx(set())
    for name, ct, init in mod.params:  # initialised at load time in declaration order (initialisers may call sqrt)
        o.append("static %s %s;" % (ct, ctxn.cname(name)))
    for name, ct, lo, vals in mod.parrays:
        o.append("static const %s %s_[] = {%s};" % (ct, name, ", ".join(vals)))
        o.append("#define %s(i) %s_[(i) - (%s)]" % (name, name, lo))
    for name, ct in mod.scalars.items():
        o.append("static %s %s;" % (ct, ctxn.cname(name)))
    for name, (ct, rank, fixed) in mod.arrays.items():
        if rank > 3:
            continue
        o.append("static %s *%s_; static ref_array_desc %s_d = {\"%s\"};" % (ct, name, name, name))
        idx = ", ".join("ijk"[:rank])
        o.append("#define %s(%s) %s_[REF_IDX%d(%s, %s)]" % (name, idx, name, rank, name, idx))
    o.append("__attribute__((constructor)) static void ref_init_parameters(void) {")
    for name, ct, init in mod.params:
        o.append("    %s = %s;" % (ctxn.cname(name), init))
    o.append("}")
    # prototypes
    for head, body, argt in subs_c.values():
        o.append(head + ";")
    for name in sorted(called - set(subs_c)):
        o.append("#define %s(...) ref_abort(\"call of %s, which was not translated\")" % (name, name))
    for name, (head, body, argt) in subs_c.items():
        o.append("\n" + head + " {\n" + body + "\n}")
    # registry
    o.append("\nstatic struct { const char *name; void *p; int kind; } ref_scalars[] = {")
    for name, ct in mod.scalars.items():
        if ct in KIND:
            o.append("    {\"%s\", &%s, %d}," % (name, ctxn.cname(name), KIND[ct]))
    o.append("    {0, 0, 0}};")
    arrs = [n for n, (ct, rank, fixed) in mod.arrays.items() if rank <= 3]
    for n in arrs:
        rank = mod.arrays[n][1]
        args = ", ".join("lo[%d], hi[%d]" % (m, m) for m in range(rank))
        o.append("static void ref_alloc_%s(int rank, const long long *lo, const long long *hi) { if (rank != %d) ref_abort(\"rank of %s\"); REF_ALLOC%d(%s, %s); }" % (n, rank, n, rank, n, args))
    o.append("static ref_array_desc *ref_arrays[] = {%s 0};" % "".join("&%s_d, " % n for n in arrs))
    o.append("static void (*ref_allocators[])(int, const long long *, const long long *) = {%s 0};" % "".join("ref_alloc_%s, " % n for n in arrs))
    for name, (head, body, argt) in subs_c.items():
        ii, di, call = 0, 0, []
        for t in argt:
            if t == "double" or t == "float":
                call.append("da[%d]" % di)
                di += 1
            else:
                call.append("(%s)ia[%d]" % (t, ii))
                ii += 1
        o.append("static void ref_thunk_%s(const long long *ia, const double *da) { (void)ia; (void)da; %s(%s); }" % (name, name, ", ".join(call)))
    o.append("static struct { const char *name; void (*fn)(const long long *, const double *); } ref_subs[] = {")
    for name in subs_c:
        o.append("    {\"%s\", ref_thunk_%s}," % (name, name))
    o.append("    {0, 0}};")
    o.append(API)
    if skipped:
        o.append("/* not translated:\n%s\n*/" % "\n".join("  %s: %s" % kv for kv in skipped.items()))
    return "\n".join(o) + "\n"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-o", "--output", required=True)
    ap.add_argument("--modules", nargs="+", required=True, help="source files holding the module declarations")
    ap.add_argument("--sources", nargs="+", required=True, help="source files holding the subroutines")
    ap.add_argument("--only", nargs="*", default=None, help="translate only these subroutines (default: every one that fits the subset)")
    ap.add_argument("--define", "-D", action="append", default=[])
    ap.add_argument("--openmp", action="store_true", help="carry '!$omp parallel do' loop directives over as '#pragma omp parallel for'")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    mod = Module()
    for p in a.modules:
        parse_modules(logical_lines(preprocess(p, a.define)), mod)
    subs_c, skipped, known = {}, {}, {}
    for p in a.sources:
        for sub in split_subroutines(logical_lines(preprocess(p, a.define))):
            if a.only is not None and sub.name not in a.only:
                continue
            try:
                called_before = set(known.get("__called__", set()))
                subs_c[sub.name] = translate_sub(sub, mod, known, a.openmp)
            except Unsupported as e:
                skipped[sub.name] = str(e)
                known["__called__"] = called_before
    called = known.get("__called__", set())
    os.makedirs(os.path.dirname(os.path.abspath(a.output)), exist_ok=True)
    with open(a.output, "w") as fh:
        fh.write(emit_c(mod, subs_c, skipped, called, [os.path.relpath(p, "/root/reference") for p in a.modules + a.sources]))
    if a.verbose:
        print("translated:", " ".join(subs_c))
        for k, v in skipped.items():
            print("skipped %s: %s" % (k, v))
    return 0


if __name__ == "__main__":
    sys.exit(main())
