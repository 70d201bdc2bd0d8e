"""minijs -- a small ECMAScript-5 subset interpreter (TEST INFRASTRUCTURE).

Why it exists: the reference (Cyame/glpk.js) is pure JavaScript and neither the
build container nor the GPU box has a JS engine, so the reference could not be
executed and parity was "unpinned".  This interpreter runs the reference's
UNMODIFIED sources where they lie (/root/reference/lib/*.js) so that golden
vectors -- results, pivot sequences and live arrays of glp_simplex / glp_intopt
-- can be generated from the reference itself (oracle/jsref/make_ref_golden.py
writes them to tests/golden/; nothing of the reference is copied into the repo).

Scope: what lib/*.js (without the MathProg translator) uses -- var/function
hoisting, closures, sloppy-mode `this`, objects, arrays, typed arrays,
switch fall-through, do/while/for/for-in, ++/--, compound assignment, the
comma/ternary/logical/bitwise operators with ToInt32 semantics, exceptions,
Math/String/Number built-ins.  No regular expressions (a literal is tokenised
and evaluates to a dummy), no getters/setters, no prototypes, no labels.

Numbers: JS has only doubles.  Integer-valued literals are kept as Python ints
(exact up to 2**53 like doubles, and usable as list indices); every operation
that can leave the integers (/, Math.*, typed float arrays) produces floats, and
division by zero, NaN truthiness, -0, % and the bitwise operators follow JS.
Compiled to Python closures with statically resolved scopes for speed.
"""
import math
import sys
import time

sys.setrecursionlimit(20000)


class _Undef:
    __slots__ = ()

    def __bool__(self):
        return False

    def __repr__(self):
        return "undefined"


UNDEF = _Undef()
NAN = float("nan")
INF = float("inf")


class JSThrow(Exception):
    def __init__(self, value):
        Exception.__init__(self, js_to_str(value.get("message", value)) if isinstance(value, dict) else js_to_str(value))
        self.value = value


class JSArray(list):
    """JS Array: grows on assignment past the end, holes read as undefined"""
    __slots__ = ()


class TypedArray(list):
    __slots__ = ()
    kind = "f64"


class Int32Array(TypedArray):
    __slots__ = ()
    kind = "i32"


class Int8Array(TypedArray):
    __slots__ = ()
    kind = "i8"


class Float64Array(TypedArray):
    __slots__ = ()
    kind = "f64"


def to_int32(v):
    if isinstance(v, int) and not isinstance(v, bool):
        if -2147483648 <= v <= 2147483647:
            return v
        return ((v + 2147483648) & 0xFFFFFFFF) - 2147483648
    v = to_num(v)
    if v != v or v in (INF, -INF):
        return 0
    return ((int(v) + 2147483648) & 0xFFFFFFFF) - 2147483648


def to_uint32(v):
    return to_int32(v) & 0xFFFFFFFF


def to_num(v):
    if isinstance(v, bool):
        return 1 if v else 0
    if isinstance(v, (int, float)):
        return v
    if v is None:
        return 0
    if v is UNDEF:
        return NAN
    if isinstance(v, str):
        s = v.strip()
        if s == "":
            return 0
        try:
            if s.startswith(("0x", "0X")):
                return int(s, 16)
            f = float(s)
            return int(f) if f == int(f) and abs(f) < 2 ** 53 and "." not in s and "e" not in s.lower() else f
        except ValueError:
            return NAN
    return NAN


def js_num_to_str(v):
    """Number::toString(10) of ECMAScript (sec. 6.1.6.1.20): the shortest round-trip digits (Python's repr
    yields the same ones) laid out by the position n of the decimal point: digits padded with zeros for
    k <= n <= 21, a point inside for 0 < n <= 21, '0.000ddd' for -6 < n <= 0, exponent form otherwise."""
    if isinstance(v, bool):
        return "true" if v else "false"
    if isinstance(v, int):
        return str(v)
    if v != v:
        return "NaN"
    if v == INF:
        return "Infinity"
    if v == -INF:
        return "-Infinity"
    if v == 0:
        return "0"
    if v < 0:
        return "-" + js_num_to_str(-v)
    r = repr(float(v))
    if "e" in r:
        mant, ex = r.split("e")
        digits, n = mant.replace(".", ""), int(ex) + 1
    else:
        ip, fp = r.split(".")
        if ip != "0":
            digits, n = ip + fp, len(ip)
        else:
            digits = fp.lstrip("0")
            n = -(len(fp) - len(digits))
    digits = digits.rstrip("0") or "0"
    k = len(digits)
    if k <= n <= 21:
        return digits + "0" * (n - k)
    if 0 < n <= 21:
        return digits[:n] + "." + digits[n:]
    if -6 < n <= 0:
        return "0." + "0" * (-n) + digits
    e = n - 1
    tail = "e" + ("+" if e > 0 else "-") + str(abs(e))
    return (digits if k == 1 else digits[0] + "." + digits[1:]) + tail


def js_to_str(v):
    if isinstance(v, str):
        return v
    if isinstance(v, (bool, int, float)):
        return js_num_to_str(v)
    if v is None:
        return "null"
    if v is UNDEF:
        return "undefined"
    if isinstance(v, list):
        return ",".join("" if (x is UNDEF or x is None) else js_to_str(x) for x in v)
    if isinstance(v, dict):
        return "[object Object]"
    return "function"


def truthy(v):
    if isinstance(v, float):
        return v == v and v != 0.0
    if isinstance(v, (dict, list)):
        return True
    return bool(v)


def js_typeof(v):
    if v is UNDEF:
        return "undefined"
    if v is None:
        return "object"
    if isinstance(v, bool):
        return "boolean"
    if isinstance(v, (int, float)):
        return "number"
    if isinstance(v, str):
        return "string"
    if isinstance(v, (JSFunc, NativeFunc)):
        return "function"
    return "object"


def loose_eq(a, b):
    if a is b:
        return not (isinstance(a, float) and a != a)
    ta, tb = type(a), type(b)
    if (a is None or a is UNDEF) or (b is None or b is UNDEF):
        return (a is None or a is UNDEF) and (b is None or b is UNDEF)
    if ta in (int, float, bool) and tb in (int, float, bool):
        return a == b
    if ta is str and tb is str:
        return a == b
    if ta in (int, float, bool) and tb is str:
        return a == to_num(b)
    if ta is str and tb in (int, float, bool):
        return to_num(a) == b
    return False


def strict_eq(a, b):
    if a is b:
        return not (isinstance(a, float) and a != a)
    ta, tb = type(a), type(b)
    if ta in (int, float) and tb in (int, float):
        return a == b
    if ta is tb and ta in (str, bool):
        return a == b
    return False


def js_add(a, b):
    if isinstance(a, str) or isinstance(b, str):
        return js_to_str(a) + js_to_str(b)
    if isinstance(a, (dict, list)) or isinstance(b, (dict, list)):
        return js_to_str(a) + js_to_str(b)
    return to_num(a) + to_num(b)


def js_div(a, b):
    a, b = to_num(a), to_num(b)
    try:
        return a / b
    except ZeroDivisionError:
        if a != a or a == 0:
            return NAN
        neg = (a < 0) != (math.copysign(1.0, b) < 0)
        return -INF if neg else INF


def js_mod(a, b):
    a, b = to_num(a), to_num(b)
    if isinstance(a, int) and isinstance(b, int):
        if b == 0:
            return NAN
        r = abs(a) % abs(b)
        return -r if a < 0 else r
    try:
        return math.fmod(a, b)
    except (ValueError, ZeroDivisionError):
        return NAN


def js_neg(a):
    a = to_num(a)
    if isinstance(a, int) and a == 0:
        return -0.0
    return -a


def js_lt(a, b):
    if isinstance(a, str) and isinstance(b, str):
        return a < b
    a, b = to_num(a), to_num(b)
    return a < b


# --------------------------------------------------------------------------
# tokenizer
# --------------------------------------------------------------------------
KEYWORDS = {"var", "function", "return", "if", "else", "for", "while", "do", "break", "continue", "switch", "case",
            "default", "new", "delete", "typeof", "in", "instanceof", "this", "null", "true", "false", "throw", "try",
            "catch", "finally", "void"}
PUNCT = [">>>=", "===", "!==", ">>>", "<<=", ">>=", "&&", "||", "==", "!=", "<=", ">=", "++", "--", "+=", "-=", "*=",
         "/=", "%=", "&=", "|=", "^=", "<<", ">>", "{", "}", "(", ")", "[", "]", ";", ",", "<", ">", "+", "-", "*", "/",
         "%", "&", "|", "^", "!", "~", "?", ":", "=", "."]


def tokenize(src, fname="<js>"):
    toks = []
    i, n, line = 0, len(src), 1
    nl = False
    idstart = set("abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ_$")
    idchar = idstart | set("0123456789")
    while i < n:
        c = src[i]
        if c == "\n":
            line += 1
            nl = True
            i += 1
            continue
        if c in " \t\r\f\v﻿":
            i += 1
            continue
        if c == "/" and i + 1 < n and src[i + 1] == "/":
            while i < n and src[i] != "\n":
                i += 1
            continue
        if c == "/" and i + 1 < n and src[i + 1] == "*":
            j = src.index("*/", i + 2)
            seg = src[i:j]
            if "\n" in seg:
                nl = True
                line += seg.count("\n")
            i = j + 2
            continue
        if c in idstart:
            j = i + 1
            while j < n and src[j] in idchar:
                j += 1
            w = src[i:j]
            toks.append(("kw" if w in KEYWORDS else "id", w, line, nl))
            i, nl = j, False
            continue
        if c.isdigit() or (c == "." and i + 1 < n and src[i + 1].isdigit()):
            j = i
            if c == "0" and i + 1 < n and src[i + 1] in "xX":
                j = i + 2
                while j < n and src[j] in "0123456789abcdefABCDEF":
                    j += 1
                toks.append(("num", int(src[i:j], 16), line, nl))
            else:
                isf = False
                while j < n and src[j].isdigit():
                    j += 1
                if j < n and src[j] == ".":
                    isf = True
                    j += 1
                    while j < n and src[j].isdigit():
                        j += 1
                if j < n and src[j] in "eE":
                    k = j + 1
                    if k < n and src[k] in "+-":
                        k += 1
                    if k < n and src[k].isdigit():
                        isf = True
                        j = k
                        while j < n and src[j].isdigit():
                            j += 1
                txt = src[i:j]
                if isf:
                    f = float(txt)
                    toks.append(("num", f, line, nl))
                else:
                    toks.append(("num", int(txt), line, nl))
            i, nl = j, False
            continue
        if c in "\"'":
            j = i + 1
            out = []
            while src[j] != c:
                ch = src[j]
                if ch == "\\":
                    j += 1
                    e = src[j]
                    if e == "n":
                        out.append("\n")
                    elif e == "t":
                        out.append("\t")
                    elif e == "r":
                        out.append("\r")
                    elif e == "v":
                        out.append("\v")
                    elif e == "f":
                        out.append("\f")
                    elif e == "b":
                        out.append("\b")
                    elif e == "0":
                        out.append("\0")
                    elif e == "x":
                        out.append(chr(int(src[j + 1:j + 3], 16)))
                        j += 2
                    elif e == "u":
                        out.append(chr(int(src[j + 1:j + 5], 16)))
                        j += 4
                    elif e == "\n":
                        line += 1
                    else:
                        out.append(e)
                else:
                    out.append(ch)
                j += 1
            toks.append(("str", "".join(out), line, nl))
            i, nl = j + 1, False
            continue
        if c == "/":
            prev = toks[-1] if toks else None
            regex_ok = prev is None or (prev[0] == "punct" and prev[1] not in (")", "]", "}")) or \
                (prev[0] == "kw" and prev[1] in ("return", "typeof", "case", "in", "new", "delete", "void", "throw"))
            if regex_ok:
                j = i + 1
                incl = False
                while True:
                    ch = src[j]
                    if ch == "\\":
                        j += 2
                        continue
                    if ch == "[":
                        incl = True
                    elif ch == "]":
                        incl = False
                    elif ch == "/" and not incl:
                        break
                    j += 1
                j += 1
                while j < n and src[j] in idchar:
                    j += 1
                toks.append(("regex", src[i:j], line, nl))
                i, nl = j, False
                continue
        for p in PUNCT:
            if src.startswith(p, i):
                toks.append(("punct", p, line, nl))
                i += len(p)
                nl = False
                break
        else:
            raise SyntaxError("%s:%d: unexpected character %r" % (fname, line, c))
    toks.append(("eof", None, line, True))
    return toks


# --------------------------------------------------------------------------
# parser -> AST (tuples)
# --------------------------------------------------------------------------
BINPREC = {"||": 1, "&&": 2, "|": 3, "^": 4, "&": 5, "==": 6, "!=": 6, "===": 6, "!==": 6, "<": 7, ">": 7, "<=": 7,
           ">=": 7, "instanceof": 7, "in": 7, "<<": 8, ">>": 8, ">>>": 8, "+": 9, "-": 9, "*": 10, "/": 10, "%": 10}
ASSIGN_OPS = {"=", "+=", "-=", "*=", "/=", "%=", "&=", "|=", "^=", "<<=", ">>=", ">>>="}


class Parser:
    def __init__(self, toks, fname):
        self.t, self.i, self.fname = toks, 0, fname
        self.no_in = False

    def peek(self):
        return self.t[self.i]

    def next(self):
        tok = self.t[self.i]
        self.i += 1
        return tok

    def err(self, msg):
        tok = self.peek()
        raise SyntaxError("%s:%d: %s (at %r)" % (self.fname, tok[2], msg, tok[1]))

    def is_p(self, v):
        tok = self.t[self.i]
        return tok[0] == "punct" and tok[1] == v

    def is_kw(self, v):
        tok = self.t[self.i]
        return tok[0] == "kw" and tok[1] == v

    def eat_p(self, v):
        if self.is_p(v):
            self.i += 1
            return True
        return False

    def expect_p(self, v):
        if not self.eat_p(v):
            self.err("expected %r" % v)

    def semi(self):
        if self.eat_p(";"):
            return
        tok = self.peek()
        if tok[0] == "eof" or (tok[0] == "punct" and tok[1] == "}") or tok[3]:
            return
        self.err("expected ';'")

    def program(self):
        body = []
        while self.peek()[0] != "eof":
            body.append(self.statement())
        return body

    def block(self):
        self.expect_p("{")
        body = []
        while not self.is_p("}"):
            body.append(self.statement())
        self.expect_p("}")
        return ("block", body)

    def function(self, is_decl):
        line = self.peek()[2]
        name = None
        if self.peek()[0] == "id":
            name = self.next()[1]
        elif is_decl:
            self.err("function name expected")
        self.expect_p("(")
        params = []
        while not self.is_p(")"):
            params.append(self.next()[1])
            if not self.eat_p(","):
                break
        self.expect_p(")")
        self.expect_p("{")
        body = []
        while not self.is_p("}"):
            body.append(self.statement())
        self.expect_p("}")
        return ("func", name, params, body, line)

    def var_decls(self):
        decls = []
        while True:
            tok = self.next()
            if tok[0] != "id":
                self.err("variable name expected")
            init = None
            if self.eat_p("="):
                init = self.assign()
            decls.append((tok[1], init))
            if not self.eat_p(","):
                break
        return ("var", decls)

    def statement(self):
        tok = self.peek()
        kind, v = tok[0], tok[1]
        if kind == "punct":
            if v == "{":
                return self.block()
            if v == ";":
                self.i += 1
                return ("empty",)
        if kind == "kw":
            if v == "var":
                self.i += 1
                d = self.var_decls()
                self.semi()
                return d
            if v == "function":
                self.i += 1
                f = self.function(True)
                return ("funcdecl", f)
            if v == "if":
                self.i += 1
                self.expect_p("(")
                c = self.expression()
                self.expect_p(")")
                a = self.statement()
                b = None
                if self.is_kw("else"):
                    self.i += 1
                    b = self.statement()
                return ("if", c, a, b)
            if v == "for":
                self.i += 1
                self.expect_p("(")
                init = None
                if self.is_kw("var"):
                    self.i += 1
                    self.no_in = True
                    init = self.var_decls()
                    self.no_in = False
                    if self.is_kw("in"):
                        self.i += 1
                        obj = self.expression()
                        self.expect_p(")")
                        return ("forin", ("id", init[1][0][0]), obj, self.statement(), init)
                elif not self.is_p(";"):
                    self.no_in = True
                    init = self.expression()
                    self.no_in = False
                    if self.is_kw("in"):
                        self.i += 1
                        obj = self.expression()
                        self.expect_p(")")
                        return ("forin", init, obj, self.statement(), None)
                    init = ("expr", init)
                self.expect_p(";")
                cond = None if self.is_p(";") else self.expression()
                self.expect_p(";")
                upd = None if self.is_p(")") else self.expression()
                self.expect_p(")")
                return ("for", init, cond, upd, self.statement())
            if v == "while":
                self.i += 1
                self.expect_p("(")
                c = self.expression()
                self.expect_p(")")
                return ("while", c, self.statement())
            if v == "do":
                self.i += 1
                body = self.statement()
                if not self.is_kw("while"):
                    self.err("expected while")
                self.i += 1
                self.expect_p("(")
                c = self.expression()
                self.expect_p(")")
                self.eat_p(";")
                return ("dowhile", body, c)
            if v == "return":
                self.i += 1
                nxt = self.peek()
                arg = None
                if not (nxt[0] == "eof" or (nxt[0] == "punct" and nxt[1] in (";", "}")) or nxt[3]):
                    arg = self.expression()
                self.semi()
                return ("return", arg)
            if v == "break":
                self.i += 1
                self.semi()
                return ("break",)
            if v == "continue":
                self.i += 1
                self.semi()
                return ("continue",)
            if v == "throw":
                self.i += 1
                e = self.expression()
                self.semi()
                return ("throw", e)
            if v == "try":
                self.i += 1
                blk = self.block()
                cname, cblk, fblk = None, None, None
                if self.is_kw("catch"):
                    self.i += 1
                    self.expect_p("(")
                    cname = self.next()[1]
                    self.expect_p(")")
                    cblk = self.block()
                if self.is_kw("finally"):
                    self.i += 1
                    fblk = self.block()
                return ("try", blk, cname, cblk, fblk)
            if v == "switch":
                self.i += 1
                self.expect_p("(")
                d = self.expression()
                self.expect_p(")")
                self.expect_p("{")
                cases = []
                while not self.is_p("}"):
                    if self.is_kw("case"):
                        self.i += 1
                        test = self.expression()
                    elif self.is_kw("default"):
                        self.i += 1
                        test = None
                    else:
                        self.err("case expected")
                    self.expect_p(":")
                    body = []
                    while not (self.is_kw("case") or self.is_kw("default") or self.is_p("}")):
                        body.append(self.statement())
                    cases.append((test, body))
                self.expect_p("}")
                return ("switch", d, cases)
        e = self.expression()
        self.semi()
        return ("expr", e)

    def expression(self):
        e = self.assign()
        while self.is_p(","):
            self.i += 1
            e = ("comma", e, self.assign())
        return e

    def assign(self):
        left = self.ternary()
        tok = self.peek()
        if tok[0] == "punct" and tok[1] in ASSIGN_OPS:
            self.i += 1
            right = self.assign()
            if left[0] not in ("id", "member", "index"):
                self.err("invalid assignment target")
            return ("assign", tok[1], left, right)
        return left

    def ternary(self):
        c = self.binary(0)
        if self.is_p("?"):
            self.i += 1
            save = self.no_in
            self.no_in = False
            a = self.assign()
            self.no_in = save
            self.expect_p(":")
            b = self.assign()
            return ("cond", c, a, b)
        return c

    def binary(self, minprec):
        left = self.unary()
        while True:
            tok = self.peek()
            op = tok[1] if tok[0] in ("punct", "kw") else None
            if op == "in" and self.no_in:
                return left
            prec = BINPREC.get(op)
            if prec is None or prec < minprec:
                return left
            self.i += 1
            right = self.binary(prec + 1)
            if op in ("&&", "||"):
                left = ("logic", op, left, right)
            else:
                left = ("bin", op, left, right)

    def unary(self):
        tok = self.peek()
        if tok[0] == "punct" and tok[1] in ("!", "-", "+", "~"):
            self.i += 1
            return ("unary", tok[1], self.unary())
        if tok[0] == "punct" and tok[1] in ("++", "--"):
            self.i += 1
            return ("update", tok[1], True, self.unary())
        if tok[0] == "kw" and tok[1] in ("typeof", "delete", "void"):
            self.i += 1
            return ("unary", tok[1], self.unary())
        e = self.postfix()
        return e

    def postfix(self):
        e = self.call_member()
        tok = self.peek()
        if tok[0] == "punct" and tok[1] in ("++", "--") and not tok[3]:
            self.i += 1
            return ("update", tok[1], False, e)
        return e

    def args(self):
        self.expect_p("(")
        a = []
        while not self.is_p(")"):
            a.append(self.assign())
            if not self.eat_p(","):
                break
        self.expect_p(")")
        return a

    def call_member(self):
        if self.is_kw("new"):
            self.i += 1
            callee = self.member_only()
            a = self.args() if self.is_p("(") else []
            e = ("new", callee, a)
        else:
            e = self.primary()
        while True:
            if self.is_p("."):
                self.i += 1
                e = ("member", e, self.next()[1])
            elif self.is_p("["):
                self.i += 1
                save = self.no_in
                self.no_in = False
                idx = self.expression()
                self.no_in = save
                self.expect_p("]")
                e = ("index", e, idx)
            elif self.is_p("("):
                e = ("call", e, self.args())
            else:
                return e

    def member_only(self):
        if self.is_kw("new"):
            self.i += 1
            callee = self.member_only()
            a = self.args() if self.is_p("(") else []
            e = ("new", callee, a)
        else:
            e = self.primary()
        while True:
            if self.is_p("."):
                self.i += 1
                e = ("member", e, self.next()[1])
            elif self.is_p("["):
                self.i += 1
                idx = self.expression()
                self.expect_p("]")
                e = ("index", e, idx)
            else:
                return e

    def primary(self):
        tok = self.next()
        kind, v = tok[0], tok[1]
        if kind == "num":
            return ("lit", v)
        if kind == "str":
            return ("lit", v)
        if kind == "regex":
            return ("lit", {"__regex__": v})
        if kind == "id":
            return ("id", v)
        if kind == "kw":
            if v == "this":
                return ("this",)
            if v == "null":
                return ("lit", None)
            if v == "true":
                return ("lit", True)
            if v == "false":
                return ("lit", False)
            if v == "function":
                return self.function(False)
        if kind == "punct":
            if v == "(":
                save = self.no_in
                self.no_in = False
                e = self.expression()
                self.no_in = save
                self.expect_p(")")
                return e
            if v == "[":
                items = []
                while not self.is_p("]"):
                    items.append(self.assign())
                    if not self.eat_p(","):
                        break
                self.expect_p("]")
                return ("array", items)
            if v == "{":
                props = []
                while not self.is_p("}"):
                    k = self.next()
                    key = js_to_str(k[1])
                    self.expect_p(":")
                    props.append((key, self.assign()))
                    if not self.eat_p(","):
                        break
                self.expect_p("}")
                return ("object", props)
        self.i -= 1
        self.err("unexpected token")


# --------------------------------------------------------------------------
# runtime objects
# --------------------------------------------------------------------------
class Env:
    __slots__ = ("v", "p", "this", "ret")

    def __init__(self, v, p, this):
        self.v, self.p, self.this, self.ret = v, p, this, UNDEF


class NativeFunc:
    __slots__ = ("fn", "name", "props")

    def __init__(self, fn, name="native"):
        self.fn, self.name, self.props = fn, name, {}

    def call(self, this, args):
        return self.fn(this, args)


class JSFunc:
    __slots__ = ("name", "qname", "params", "body", "env", "decl_names", "interp", "props", "nparams", "uses_args")

    def call(self, this, args):
        v = dict.fromkeys(self.decl_names, UNDEF)
        np_ = self.nparams
        na = len(args)
        params = self.params
        if na >= np_:
            for i in range(np_):
                v[params[i]] = args[i]
        else:
            for i in range(na):
                v[params[i]] = args[i]
        if self.uses_args:
            v["arguments"] = JSArray(args)
        env = Env(v, self.env, this)
        hooks = self.interp.hooks
        if hooks:
            h = hooks.get(self.qname)
            if h is not None:
                h("enter", self.qname, args, None)
                self.body(env)
                h("exit", self.qname, args, env.ret)
                return env.ret
        self.body(env)
        return env.ret


NORMAL, BREAK, CONTINUE, RETURN = 0, 1, 2, 3


def get_prop(obj, key, interp):
    """obj.key / obj[key]"""
    if isinstance(obj, dict):
        if isinstance(key, str):
            return obj.get(key, UNDEF)
        return obj.get(js_to_str(key), UNDEF)
    if isinstance(obj, list):
        if isinstance(key, int) and not isinstance(key, bool):
            if 0 <= key < len(obj):
                return obj[key]
            return UNDEF
        if isinstance(key, float):
            if key == int(key):
                k = int(key)
                if 0 <= k < len(obj):
                    return obj[k]
            return UNDEF
        if key == "length":
            return len(obj)
        m = ARRAY_METHODS.get(key)
        if m is not None:
            return BoundMethod(obj, m)
        if isinstance(key, str) and key.isdigit():
            return get_prop(obj, int(key), interp)
        return UNDEF
    if isinstance(obj, str):
        if key == "length":
            return len(obj)
        if isinstance(key, int):
            return obj[key] if 0 <= key < len(obj) else UNDEF
        m = STRING_METHODS.get(key)
        if m is not None:
            return BoundMethod(obj, m)
        return UNDEF
    if isinstance(obj, (int, float)) and not isinstance(obj, bool):
        m = NUMBER_METHODS.get(key)
        if m is not None:
            return BoundMethod(obj, m)
        return UNDEF
    if isinstance(obj, (JSFunc, NativeFunc)):
        return obj.props.get(key, UNDEF)
    if obj is UNDEF or obj is None:
        raise JSThrow({"message": "TypeError: cannot read property %r of %s" % (key, js_to_str(obj))})
    return UNDEF


def set_prop(obj, key, val):
    if isinstance(obj, dict):
        obj[key if isinstance(key, str) else js_to_str(key)] = val
        return
    if isinstance(obj, list):
        if isinstance(key, float) and key == int(key):
            key = int(key)
        if isinstance(key, int):
            if isinstance(obj, TypedArray):
                if 0 <= key < len(obj):
                    k = obj.kind
                    if k == "f64":
                        obj[key] = float(to_num(val))
                    elif k == "i32":
                        obj[key] = to_int32(val)
                    else:
                        v = to_int32(val) & 0xFF
                        obj[key] = v - 256 if v >= 128 else v
                return
            n = len(obj)
            if key < n:
                obj[key] = val
            else:
                if key > n:
                    obj.extend([UNDEF] * (key - n))
                obj.append(val)
            return
        if key == "length":
            del obj[int(val):]
            return
        return
    if isinstance(obj, (JSFunc, NativeFunc)):
        obj.props[key] = val
        return
    if obj is UNDEF or obj is None:
        raise JSThrow({"message": "TypeError: cannot set property %r of %s" % (key, js_to_str(obj))})


class BoundMethod:
    __slots__ = ("obj", "fn")

    def __init__(self, obj, fn):
        self.obj, self.fn = obj, fn

    def call(self, this, args):
        return self.fn(self.obj, args)


def _arr_sort(a, args):
    import functools
    if args and args[0] is not UNDEF:
        f = args[0]
        a.sort(key=functools.cmp_to_key(lambda x, y: (lambda r: -1 if r < 0 else (1 if r > 0 else 0))(to_num(f.call(UNDEF, [x, y])))))
    else:
        a.sort(key=js_to_str)
    return a


def _arr_slice(a, args):
    n = len(a)
    s = int(to_num(args[0])) if args else 0
    e = int(to_num(args[1])) if len(args) > 1 and args[1] is not UNDEF else n
    if s < 0:
        s = max(0, n + s)
    if e < 0:
        e = max(0, n + e)
    return JSArray(a[s:e])


ARRAY_METHODS = {
    "push": lambda a, args: (a.extend(args), len(a))[1],
    "pop": lambda a, args: a.pop() if a else UNDEF,
    "join": lambda a, args: (js_to_str(args[0]) if args else ",").join("" if (x is UNDEF or x is None) else js_to_str(x) for x in a),
    "sort": _arr_sort,
    "slice": _arr_slice,
    "indexOf": lambda a, args: next((i for i, x in enumerate(a) if strict_eq(x, args[0])), -1),
    "concat": lambda a, args: JSArray(list(a) + [y for x in args for y in (x if isinstance(x, list) else [x])]),
}


def _str_slice(s, args):
    n = len(s)
    b = int(to_num(args[0])) if args else 0
    e = int(to_num(args[1])) if len(args) > 1 and args[1] is not UNDEF else n
    if b < 0:
        b = max(0, n + b)
    if e < 0:
        e = max(0, n + e)
    return s[b:e]


def _str_substr(s, args):
    b = int(to_num(args[0])) if args else 0
    if b < 0:
        b = max(0, len(s) + b)
    ln = int(to_num(args[1])) if len(args) > 1 and args[1] is not UNDEF else len(s)
    return s[b:b + ln]


def _str_substring(s, args):
    n = len(s)
    b = min(max(int(to_num(args[0])) if args else 0, 0), n)
    e = min(max(int(to_num(args[1])), 0), n) if len(args) > 1 and args[1] is not UNDEF else n
    if b > e:
        b, e = e, b
    return s[b:e]


STRING_METHODS = {
    "charAt": lambda s, args: (lambda i: s[i] if 0 <= i < len(s) else "")(int(to_num(args[0])) if args else 0),
    "charCodeAt": lambda s, args: (lambda i: ord(s[i]) if 0 <= i < len(s) else NAN)(int(to_num(args[0])) if args else 0),
    "indexOf": lambda s, args: s.find(js_to_str(args[0]), int(to_num(args[1])) if len(args) > 1 else 0),
    "lastIndexOf": lambda s, args: s.rfind(js_to_str(args[0])),
    "toLowerCase": lambda s, args: s.lower(),
    "toUpperCase": lambda s, args: s.upper(),
    "slice": _str_slice,
    "substr": _str_substr,
    "substring": _str_substring,
    "toString": lambda s, args: s,
    "split": lambda s, args: JSArray(s.split(js_to_str(args[0])) if args and args[0] != "" else list(s)),
    "replace": lambda s, args: s.replace(js_to_str(args[0]), js_to_str(args[1]), 1) if isinstance(args[0], str) else s,
    "trim": lambda s, args: s.strip(),
}


def _num_tofixed(v, args):
    d = int(to_num(args[0])) if args else 0
    return "%.*f" % (d, float(v))


def _num_tostring(v, args):
    if args and args[0] is not UNDEF and int(to_num(args[0])) != 10:
        base = int(to_num(args[0]))
        n = int(v)
        digs = "0123456789abcdefghijklmnopqrstuvwxyz"
        if n == 0:
            return "0"
        neg, n = n < 0, abs(n)
        out = ""
        while n:
            out = digs[n % base] + out
            n //= base
        return ("-" if neg else "") + out
    return js_num_to_str(v)


NUMBER_METHODS = {"toFixed": _num_tofixed, "toString": _num_tostring,
                  "toPrecision": lambda v, args: "%.*g" % (int(to_num(args[0])), float(v)),
                  "toExponential": lambda v, args: "%.*e" % (int(to_num(args[0])) if args else 6, float(v))}


# --------------------------------------------------------------------------
# compiler: AST -> closures
# --------------------------------------------------------------------------
class Scope:
    def __init__(self, names, parent, fname):
        self.names, self.parent, self.fname = names, parent, fname


def collect_decls(body, names):
    """var and function declarations of one function body (not nested functions)"""
    for st in body:
        _collect_stmt(st, names)


def _collect_stmt(st, names):
    k = st[0]
    if k == "var":
        for n, _ in st[1]:
            names.add(n)
    elif k == "funcdecl":
        names.add(st[1][1])
    elif k == "block":
        collect_decls(st[1], names)
    elif k == "if":
        _collect_stmt(st[2], names)
        if st[3]:
            _collect_stmt(st[3], names)
    elif k == "for":
        if st[1]:
            _collect_stmt(st[1], names)
        _collect_stmt(st[4], names)
    elif k == "forin":
        if st[4]:
            _collect_stmt(st[4], names)
        _collect_stmt(st[3], names)
    elif k in ("while",):
        _collect_stmt(st[2], names)
    elif k == "dowhile":
        _collect_stmt(st[1], names)
    elif k == "try":
        _collect_stmt(st[1], names)
        if st[3]:
            _collect_stmt(st[3], names)
            names.add(st[2])
        if st[4]:
            _collect_stmt(st[4], names)
    elif k == "switch":
        for _, body in st[2]:
            collect_decls(body, names)


def _uses_arguments(node):
    if isinstance(node, tuple):
        if node and node[0] == "id" and node[1] == "arguments":
            return True
        if node and node[0] == "func":
            return False
        return any(_uses_arguments(x) for x in node)
    if isinstance(node, list):
        return any(_uses_arguments(x) for x in node)
    return False


class Interp:
    def __init__(self):
        self.globals = {}
        self.genv = Env(self.globals, None, self.globals)
        self.hooks = {}
        self.prints = []
        self.steps = 0
        install_builtins(self)

    # ---- public ----
    def run(self, src, fname="<js>"):
        ast = Parser(tokenize(src, fname), fname).program()
        names = set()
        collect_decls(ast, names)
        for n in names:
            self.globals.setdefault(n, UNDEF)
        scope = Scope(None, None, "")          # names None = global: everything late-bound
        body = self.c_body(ast, scope)
        body(self.genv)

    def call(self, f, *args):
        return f.call(self.globals, list(args))

    def get(self, name):
        return self.globals.get(name, UNDEF)

    # ---- scope resolution ----
    def resolve(self, name, scope):
        hops = 0
        s = scope
        while s is not None and s.names is not None:
            if name in s.names:
                return hops
            s = s.parent
            hops += 1
        return -1       # global

    # ---- statements ----
    def c_body(self, stmts, scope):
        """hoist function declarations, then run the statements"""
        funcs = [(st[1][1], self.c_func(st[1], scope)) for st in self._func_decls(stmts)]
        code = [self.c_stmt(st, scope) for st in stmts if st[0] not in ("funcdecl", "empty")]
        hoisted = [(self.c_store(("id", n), scope), f) for n, f in funcs]

        def run(env):
            for store, f in hoisted:
                store(env, f(env))
            for c in code:
                r = c(env)
                if r:
                    return r
            return 0
        return run

    def _func_decls(self, stmts):
        """function declarations anywhere in the body (not inside nested functions):
        ES5-era engines hoist declarations inside blocks to the function's top"""
        out = []
        for st in stmts:
            if st is None:
                continue
            k = st[0]
            if k == "funcdecl":
                out.append(st)
            elif k == "block":
                out.extend(self._func_decls(st[1]))
            elif k == "if":
                out.extend(self._func_decls([st[2], st[3]]))
            elif k == "for":
                out.extend(self._func_decls([st[4]]))
            elif k == "forin":
                out.extend(self._func_decls([st[3]]))
            elif k == "while":
                out.extend(self._func_decls([st[2]]))
            elif k == "dowhile":
                out.extend(self._func_decls([st[1]]))
            elif k == "try":
                out.extend(self._func_decls([st[1], st[3], st[4]]))
            elif k == "switch":
                for _, body in st[2]:
                    out.extend(self._func_decls(body))
        return out

    def c_block(self, stmts, scope):
        code = [self.c_stmt(st, scope) for st in stmts if st[0] not in ("funcdecl", "empty")]
        if len(code) == 1:
            return code[0]

        def run(env):
            for c in code:
                r = c(env)
                if r:
                    return r
            return 0
        return run

    def c_stmt(self, st, scope):
        k = st[0]
        if k == "expr":
            e = self.c_expr(st[1], scope)

            def run(env):
                e(env)
                return 0
            return run
        if k == "var":
            parts = [(self.c_store(("id", n), scope), self.c_expr(init, scope)) for n, init in st[1] if init is not None]
            if not parts:
                return lambda env: 0
            if len(parts) == 1:
                store, init = parts[0]

                def run1(env):
                    store(env, init(env))
                    return 0
                return run1

            def run(env):
                for store, init in parts:
                    store(env, init(env))
                return 0
            return run
        if k == "block":
            return self.c_block(st[1], scope)
        if k == "if":
            c = self.c_expr(st[1], scope)
            a = self.c_stmt(st[2], scope)
            b = self.c_stmt(st[3], scope) if st[3] else None
            if b is None:
                def run(env):
                    if truthy(c(env)):
                        return a(env)
                    return 0
            else:
                def run(env):
                    if truthy(c(env)):
                        return a(env)
                    return b(env)
            return run
        if k == "for":
            init = self.c_stmt(st[1], scope) if st[1] else None
            cond = self.c_expr(st[2], scope) if st[2] else None
            upd = self.c_expr(st[3], scope) if st[3] else None
            body = self.c_stmt(st[4], scope)

            def run(env):
                if init:
                    init(env)
                while cond is None or truthy(cond(env)):
                    r = body(env)
                    if r:
                        if r == 1:
                            break
                        if r == 3:
                            return 3
                    if upd:
                        upd(env)
                return 0
            return run
        if k == "while":
            cond = self.c_expr(st[1], scope)
            body = self.c_stmt(st[2], scope)

            def run(env):
                while truthy(cond(env)):
                    r = body(env)
                    if r:
                        if r == 1:
                            break
                        if r == 3:
                            return 3
                return 0
            return run
        if k == "dowhile":
            body = self.c_stmt(st[1], scope)
            cond = self.c_expr(st[2], scope)

            def run(env):
                while True:
                    r = body(env)
                    if r:
                        if r == 1:
                            break
                        if r == 3:
                            return 3
                    if not truthy(cond(env)):
                        break
                return 0
            return run
        if k == "forin":
            store = self.c_store(st[1], scope)
            obj = self.c_expr(st[2], scope)
            body = self.c_stmt(st[3], scope)

            def run(env):
                o = obj(env)
                if isinstance(o, dict):
                    keys = list(o.keys())
                elif isinstance(o, list):
                    keys = [str(i) for i in range(len(o))]
                else:
                    keys = []
                for key in keys:
                    store(env, key)
                    r = body(env)
                    if r:
                        if r == 1:
                            break
                        if r == 3:
                            return 3
                return 0
            return run
        if k == "return":
            e = self.c_expr(st[1], scope) if st[1] else None
            if e is None:
                def run(env):
                    env.ret = UNDEF
                    return 3
            else:
                def run(env):
                    env.ret = e(env)
                    return 3
            return run
        if k == "break":
            return lambda env: 1
        if k == "continue":
            return lambda env: 2
        if k == "throw":
            e = self.c_expr(st[1], scope)

            def run(env):
                raise JSThrow(e(env))
            return run
        if k == "try":
            blk = self.c_stmt(st[1], scope)
            cstore = self.c_store(("id", st[2]), scope) if st[3] else None
            cblk = self.c_stmt(st[3], scope) if st[3] else None
            fblk = self.c_stmt(st[4], scope) if st[4] else None

            def run(env):
                try:
                    try:
                        r = blk(env)
                    except JSThrow as ex:
                        if cblk is None:
                            raise
                        cstore(env, ex.value)
                        r = cblk(env)
                finally:
                    if fblk:
                        fr = fblk(env)
                        if fr:
                            return fr
                return r
            return run
        if k == "switch":
            disc = self.c_expr(st[1], scope)
            tests = [self.c_expr(t, scope) if t is not None else None for t, _ in st[2]]
            bodies = [self.c_block(b, scope) if b else None for _, b in st[2]]
            ncase = len(tests)
            default_ix = next((i for i, t in enumerate(tests) if t is None), -1)
            # constant cases (the usual shape): precompute a jump table
            const_tab = None
            if all(t is None or st[2][i][0][0] in ("lit", "id") for i, t in enumerate(tests)):
                const_tab = {}

            def run(env):
                d = disc(env)
                start = -1
                for i in range(ncase):
                    t = tests[i]
                    if t is not None and strict_eq(d, t(env)):
                        start = i
                        break
                if start < 0:
                    start = default_ix
                    if start < 0:
                        return 0
                for i in range(start, ncase):
                    b = bodies[i]
                    if b is None:
                        continue
                    r = b(env)
                    if r:
                        if r == 1:
                            return 0
                        return r
                return 0
            return run
        if k == "funcdecl":
            return lambda env: 0
        if k == "empty":
            return lambda env: 0
        raise SyntaxError("statement kind %r" % (k,))

    # ---- functions ----
    def c_func(self, node, scope):
        _, name, params, body, line = node
        names = set(params)
        collect_decls(body, names)
        uses_args = _uses_arguments(body)
        if uses_args:
            names.add("arguments")
        fexpr_self = name is not None
        qname = (scope.fname + "." if scope.fname else "") + (name or "<anon>")
        fscope = Scope(names, scope, qname)
        code = self.c_body(body, fscope)
        decl_names = tuple(names)
        interp = self

        def make(env):
            f = JSFunc()
            f.name, f.qname, f.params, f.nparams = name, qname, params, len(params)
            f.body, f.env, f.decl_names, f.interp, f.props, f.uses_args = code, env, decl_names, interp, {}, uses_args
            return f
        return make

    # ---- expressions ----
    def c_load(self, name, scope):
        hops = self.resolve(name, scope)
        g = self.globals
        if hops == 0:
            return lambda env: env.v[name]
        if hops == 1:
            return lambda env: env.p.v[name]
        if hops == 2:
            return lambda env: env.p.p.v[name]
        if hops == 3:
            return lambda env: env.p.p.p.v[name]
        if hops > 3:
            def load(env):
                e = env
                for _ in range(hops):
                    e = e.p
                return e.v[name]
            return load

        def gload(env):
            try:
                return g[name]
            except KeyError:
                raise JSThrow({"message": "ReferenceError: %s is not defined" % name})
        return gload

    def c_store(self, target, scope):
        """returns store(env, value)"""
        k = target[0]
        if k == "id":
            name = target[1]
            hops = self.resolve(name, scope)
            g = self.globals
            if hops == 0:
                def st0(env, v):
                    env.v[name] = v
                return st0
            if hops == 1:
                def st1(env, v):
                    env.p.v[name] = v
                return st1
            if hops > 1:
                def stn(env, v):
                    e = env
                    for _ in range(hops):
                        e = e.p
                    e.v[name] = v
                return stn

            def stg(env, v):
                g[name] = v
            return stg
        if k == "member":
            obj = self.c_expr(target[1], scope)
            key = target[2]

            def stm(env, v):
                o = obj(env)
                if type(o) is dict:
                    o[key] = v
                else:
                    set_prop(o, key, v)
            return stm
        if k == "index":
            obj = self.c_expr(target[1], scope)
            idx = self.c_expr(target[2], scope)

            def sti(env, v):
                o = obj(env)
                i = idx(env)
                if type(o) is Float64Array and type(i) is int and 0 <= i < len(o):
                    o[i] = float(v) if type(v) in (int, float) else float(to_num(v))
                else:
                    set_prop(o, i, v)
            return sti
        raise SyntaxError("bad assignment target")

    def c_expr(self, e, scope):
        k = e[0]
        if k == "lit":
            v = e[1]
            if isinstance(v, dict):
                return lambda env: dict(v)
            return lambda env: v
        if k == "id":
            name = e[1]
            if name == "undefined":
                return lambda env: UNDEF
            return self.c_load(name, scope)
        if k == "this":
            return lambda env: env.this
        if k == "member":
            obj = self.c_expr(e[1], scope)
            key = e[2]
            interp = self

            def member(env):
                o = obj(env)
                if type(o) is dict:
                    return o.get(key, UNDEF)
                return get_prop(o, key, interp)
            return member
        if k == "index":
            obj = self.c_expr(e[1], scope)
            idx = self.c_expr(e[2], scope)
            interp = self

            def index(env):
                o = obj(env)
                i = idx(env)
                if type(i) is int and isinstance(o, list):
                    if 0 <= i < len(o):
                        return o[i]
                    return UNDEF
                return get_prop(o, i, interp)
            return index
        if k == "call":
            return self.c_call(e, scope)
        if k == "new":
            callee = self.c_expr(e[1], scope)
            args = [self.c_expr(a, scope) for a in e[2]]

            def new(env):
                f = callee(env)
                a = [x(env) for x in args]
                if isinstance(f, NativeFunc):
                    return f.fn(None, a)
                if not isinstance(f, JSFunc):
                    raise JSThrow({"message": "TypeError: not a constructor"})
                o = {}
                r = f.call(o, a)
                return r if isinstance(r, (dict, list)) else o
            return new
        if k == "func":
            mk = self.c_func(e, scope)
            return mk
        if k == "assign":
            return self.c_assign(e, scope)
        if k == "update":
            return self.c_update(e, scope)
        if k == "cond":
            c, a, b = self.c_expr(e[1], scope), self.c_expr(e[2], scope), self.c_expr(e[3], scope)
            return lambda env: a(env) if truthy(c(env)) else b(env)
        if k == "comma":
            a, b = self.c_expr(e[1], scope), self.c_expr(e[2], scope)

            def comma(env):
                a(env)
                return b(env)
            return comma
        if k == "logic":
            a, b = self.c_expr(e[2], scope), self.c_expr(e[3], scope)
            if e[1] == "&&":
                def land(env):
                    v = a(env)
                    return b(env) if truthy(v) else v
                return land

            def lor(env):
                v = a(env)
                return v if truthy(v) else b(env)
            return lor
        if k == "unary":
            op = e[1]
            if op == "typeof" and e[2][0] == "id":
                name = e[2][1]
                hops = self.resolve(name, scope)
                if hops < 0:
                    g = self.globals
                    return lambda env: js_typeof(g.get(name, UNDEF))
            if op == "delete":
                t = e[2]
                if t[0] == "member":
                    obj, key = self.c_expr(t[1], scope), t[2]

                    def dele(env):
                        o = obj(env)
                        if isinstance(o, dict):
                            o.pop(key, None)
                        return True
                    return dele
                if t[0] == "index":
                    obj, idx = self.c_expr(t[1], scope), self.c_expr(t[2], scope)

                    def deli(env):
                        o = obj(env)
                        i = idx(env)
                        if isinstance(o, dict):
                            o.pop(i if isinstance(i, str) else js_to_str(i), None)
                        elif isinstance(o, list) and isinstance(i, int) and 0 <= i < len(o):
                            o[i] = UNDEF
                        return True
                    return deli
                return lambda env: True
            a = self.c_expr(e[2], scope)
            if op == "!":
                return lambda env: not truthy(a(env))
            if op == "-":
                def neg(env):
                    v = a(env)
                    if type(v) is float:
                        return -v
                    return js_neg(v)
                return neg
            if op == "+":
                return lambda env: to_num(a(env))
            if op == "~":
                return lambda env: ~to_int32(a(env))
            if op == "typeof":
                return lambda env: js_typeof(a(env))
            if op == "void":
                def void(env):
                    a(env)
                    return UNDEF
                return void
        if k == "bin":
            return self.c_bin(e, scope)
        if k == "array":
            items = [self.c_expr(x, scope) for x in e[1]]
            return lambda env: JSArray(x(env) for x in items)
        if k == "object":
            props = [(key, self.c_expr(v, scope)) for key, v in e[1]]
            return lambda env: {key: v(env) for key, v in props}
        raise SyntaxError("expression kind %r" % (k,))

    def c_call(self, e, scope):
        callee = e[1]
        args = [self.c_expr(a, scope) for a in e[2]]
        nargs = len(args)
        interp = self
        g = self.globals
        if callee[0] in ("member", "index"):
            obj = self.c_expr(callee[1], scope)
            key = callee[2] if callee[0] == "member" else None
            idx = self.c_expr(callee[2], scope) if callee[0] == "index" else None

            def mcall(env):
                o = obj(env)
                kk = key if idx is None else idx(env)
                if type(o) is dict:
                    f = o.get(kk if isinstance(kk, str) else js_to_str(kk), UNDEF)
                else:
                    f = get_prop(o, kk, interp)
                a = [x(env) for x in args]
                try:
                    return f.call(o, a)
                except AttributeError:
                    if not hasattr(f, "call"):
                        raise JSThrow({"message": "TypeError: %s is not a function" % (kk,)})
                    raise
            return mcall
        f_ = self.c_expr(callee, scope)
        cname = callee[1] if callee[0] == "id" else "<expr>"
        if nargs == 0:
            def call0(env):
                f = f_(env)
                try:
                    return f.call(g, [])
                except AttributeError:
                    if not hasattr(f, "call"):
                        raise JSThrow({"message": "TypeError: %s is not a function" % cname})
                    raise
            return call0

        def call(env):
            f = f_(env)
            a = [x(env) for x in args]
            try:
                return f.call(g, a)
            except AttributeError:
                if not hasattr(f, "call"):
                    raise JSThrow({"message": "TypeError: %s is not a function" % cname})
                raise
        return call

    def c_assign(self, e, scope):
        _, op, target, rhs = e
        r = self.c_expr(rhs, scope)
        store = self.c_store(target, scope)
        if op == "=":
            if target[0] == "id":
                def assign_id(env):
                    v = r(env)
                    store(env, v)
                    return v
                return assign_id
            if target[0] == "index":
                obj = self.c_expr(target[1], scope)
                idx = self.c_expr(target[2], scope)

                def assign_ix(env):
                    o = obj(env)
                    i = idx(env)
                    v = r(env)
                    if type(o) is Float64Array and type(i) is int and 0 <= i < len(o):
                        o[i] = float(v) if type(v) in (int, float) else float(to_num(v))
                    else:
                        set_prop(o, i, v)
                    return v
                return assign_ix
            obj = self.c_expr(target[1], scope)
            key = target[2]

            def assign_m(env):
                o = obj(env)
                v = r(env)
                if type(o) is dict:
                    o[key] = v
                else:
                    set_prop(o, key, v)
                return v
            return assign_m
        binop = BINOPS[op[:-1]]
        if target[0] == "id":
            load = self.c_load(target[1], scope)

            def cassign_id(env):
                v = binop(load(env), r(env))
                store(env, v)
                return v
            return cassign_id
        obj = self.c_expr(target[1], scope)
        idx = self.c_expr(target[2], scope) if target[0] == "index" else None
        key = target[2] if target[0] == "member" else None
        interp = self

        def cassign(env):
            o = obj(env)
            kk = key if idx is None else idx(env)
            old = get_prop(o, kk, interp)
            v = binop(old, r(env))
            set_prop(o, kk, v)
            return v
        return cassign

    def c_update(self, e, scope):
        _, op, prefix, target = e
        d = 1 if op == "++" else -1
        interp = self
        if target[0] == "id":
            load = self.c_load(target[1], scope)
            store = self.c_store(target, scope)

            def upd_id(env):
                old = load(env)
                if type(old) is not int:
                    old = to_num(old)
                new = old + d
                store(env, new)
                return new if prefix else old
            return upd_id
        obj = self.c_expr(target[1], scope)
        idx = self.c_expr(target[2], scope) if target[0] == "index" else None
        key = target[2] if target[0] == "member" else None

        def upd(env):
            o = obj(env)
            kk = key if idx is None else idx(env)
            old = to_num(get_prop(o, kk, interp))
            new = old + d
            set_prop(o, kk, new)
            return new if prefix else old
        return upd

    def c_bin(self, e, scope):
        _, op, l, r = e
        a, b = self.c_expr(l, scope), self.c_expr(r, scope)
        num = (int, float)
        if op == "+":
            def add(env):
                x, y = a(env), b(env)
                tx, ty = type(x), type(y)
                if (tx is float or tx is int) and (ty is float or ty is int):
                    return x + y
                return js_add(x, y)
            return add
        if op == "-":
            def sub(env):
                x, y = a(env), b(env)
                tx, ty = type(x), type(y)
                if (tx is float or tx is int) and (ty is float or ty is int):
                    return x - y
                return to_num(x) - to_num(y)
            return sub
        if op == "*":
            def mul(env):
                x, y = a(env), b(env)
                tx, ty = type(x), type(y)
                if tx is float or ty is float:
                    if (tx is float or tx is int) and (ty is float or ty is int):
                        return x * y
                elif tx is int and ty is int:
                    v = x * y
                    if -9007199254740992 <= v <= 9007199254740992:
                        if v == 0 and (x < 0) != (y < 0) and (x != 0 or y != 0):
                            return -0.0
                        return v
                    return float(x) * float(y)
                return to_num(x) * to_num(y)
            return mul
        if op == "/":
            def div(env):
                x, y = a(env), b(env)
                if type(y) is float and y != 0.0 and type(x) in num:
                    return x / y
                if type(y) is int and y != 0 and type(x) in num:
                    return x / y
                return js_div(x, y)
            return div
        if op == "<":
            def lt(env):
                x, y = a(env), b(env)
                if type(x) in num and type(y) in num:
                    return x < y
                return js_lt(x, y)
            return lt
        if op == ">":
            def gt(env):
                x, y = a(env), b(env)
                if type(x) in num and type(y) in num:
                    return x > y
                return js_lt(y, x)
            return gt
        if op == "<=":
            def le(env):
                x, y = a(env), b(env)
                if type(x) in num and type(y) in num:
                    return x <= y
                return BINOPS["<="](x, y)
            return le
        if op == ">=":
            def ge(env):
                x, y = a(env), b(env)
                if type(x) in num and type(y) in num:
                    return x >= y
                return BINOPS[">="](x, y)
            return ge
        if op == "==":
            def eq(env):
                x, y = a(env), b(env)
                if type(x) in num and type(y) in num:
                    return x == y
                return loose_eq(x, y)
            return eq
        if op == "!=":
            def ne(env):
                x, y = a(env), b(env)
                if type(x) in num and type(y) in num:
                    return x != y
                return not loose_eq(x, y)
            return ne
        f = BINOPS[op]
        return lambda env: f(a(env), b(env))


def _cmp_le(x, y):
    if isinstance(x, str) and isinstance(y, str):
        return x <= y
    x, y = to_num(x), to_num(y)
    return x <= y


def _js_in(key, obj):
    if isinstance(obj, dict):
        return (key if isinstance(key, str) else js_to_str(key)) in obj
    if isinstance(obj, list):
        try:
            i = int(to_num(key))
        except (ValueError, OverflowError):
            return False
        return 0 <= i < len(obj)
    return False


BINOPS = {
    "+": js_add,
    "-": lambda x, y: to_num(x) - to_num(y),
    "*": lambda x, y: to_num(x) * to_num(y),
    "/": js_div,
    "%": js_mod,
    "<": js_lt,
    ">": lambda x, y: js_lt(y, x),
    "<=": _cmp_le,
    ">=": lambda x, y: _cmp_le(y, x),
    "==": loose_eq,
    "!=": lambda x, y: not loose_eq(x, y),
    "===": strict_eq,
    "!==": lambda x, y: not strict_eq(x, y),
    "&": lambda x, y: to_int32(x) & to_int32(y),
    "|": lambda x, y: to_int32(x) | to_int32(y),
    "^": lambda x, y: to_int32(x) ^ to_int32(y),
    "<<": lambda x, y: to_int32(to_int32(x) << (to_uint32(y) & 31)),
    ">>": lambda x, y: to_int32(x) >> (to_uint32(y) & 31),
    ">>>": lambda x, y: to_uint32(x) >> (to_uint32(y) & 31),
    "in": _js_in,
    "instanceof": lambda x, y: False,
}


# --------------------------------------------------------------------------
# built-ins
# --------------------------------------------------------------------------
def install_builtins(I):
    g = I.globals

    def nf(fn, name):
        return NativeFunc(lambda this, args: fn(*args), name)

    def num1(fn):
        def w(this, args):
            x = to_num(args[0]) if args else NAN
            try:
                return fn(x)
            except (ValueError, OverflowError):
                return NAN
        return w

    def m_floor(x):
        if isinstance(x, int):
            return x
        if x != x or x in (INF, -INF):
            return x
        return float(math.floor(x))

    def m_ceil(x):
        if isinstance(x, int):
            return x
        if x != x or x in (INF, -INF):
            return x
        r = float(math.ceil(x))
        return -0.0 if r == 0 and x < 0 else r

    def m_round(x):
        if isinstance(x, int):
            return x
        if x != x or x in (INF, -INF):
            return x
        return float(math.floor(x + 0.5))

    def m_log(x):
        if x == 0:
            return -INF
        if x < 0:
            return NAN
        return math.log(x)

    def m_pow(this, args):
        x, y = to_num(args[0]), to_num(args[1])
        try:
            r = math.pow(x, y)
            if isinstance(x, int) and isinstance(y, int) and y >= 0 and abs(r) < 2 ** 53:
                return int(r)
            return r
        except OverflowError:
            return INF
        except (ValueError, ZeroDivisionError):
            return NAN

    def m_minmax(fn, empty):
        def w(this, args):
            if not args:
                return empty
            vals = [to_num(a) for a in args]
            if any(v != v for v in vals):
                return NAN
            return fn(vals)
        return w

    Math = {
        "abs": NativeFunc(num1(abs), "abs"), "floor": NativeFunc(num1(m_floor), "floor"),
        "ceil": NativeFunc(num1(m_ceil), "ceil"), "round": NativeFunc(num1(m_round), "round"),
        "sqrt": NativeFunc(num1(lambda x: math.sqrt(x) if x >= 0 else NAN), "sqrt"),
        "log": NativeFunc(num1(m_log), "log"), "exp": NativeFunc(num1(math.exp), "exp"),
        "sin": NativeFunc(num1(math.sin), "sin"), "cos": NativeFunc(num1(math.cos), "cos"),
        "atan": NativeFunc(num1(math.atan), "atan"),
        "atan2": NativeFunc(lambda this, a: math.atan2(to_num(a[0]), to_num(a[1])), "atan2"),
        "pow": NativeFunc(m_pow, "pow"),
        "min": NativeFunc(m_minmax(min, INF), "min"), "max": NativeFunc(m_minmax(max, -INF), "max"),
        "random": NativeFunc(lambda this, a: 0.5, "random"),
        "LN10": math.log(10.0), "LN2": math.log(2.0), "PI": math.pi, "E": math.e,
    }
    g["Math"] = Math
    number = NativeFunc(lambda this, a: to_num(a[0]) if a else 0, "Number")
    number.props.update({"MAX_VALUE": sys.float_info.max, "MIN_VALUE": 5e-324, "POSITIVE_INFINITY": INF,
                         "NEGATIVE_INFINITY": -INF, "NaN": NAN})
    g["Number"] = number
    string = NativeFunc(lambda this, a: js_to_str(a[0]) if a else "", "String")
    string.props["fromCharCode"] = NativeFunc(lambda this, a: "".join(chr(int(to_num(x))) for x in a), "fromCharCode")
    g["String"] = string
    g["NaN"], g["Infinity"] = NAN, INF

    def mk_typed(cls, zero):
        def ctor(this, a):
            x = a[0] if a else 0
            if isinstance(x, list):
                out = cls([zero] * len(x))
                for i, v in enumerate(x):
                    set_prop(out, i, v)
                return out
            return cls([zero] * int(to_num(x)))
        return NativeFunc(ctor, cls.__name__)
    g["Int32Array"] = mk_typed(Int32Array, 0)
    g["Int8Array"] = mk_typed(Int8Array, 0)
    g["Float64Array"] = mk_typed(Float64Array, 0.0)

    def array_ctor(this, a):
        if len(a) == 1 and isinstance(a[0], (int, float)) and not isinstance(a[0], bool):
            return JSArray([UNDEF] * int(a[0]))
        return JSArray(a)
    g["Array"] = NativeFunc(array_ctor, "Array")
    obj = NativeFunc(lambda this, a: {}, "Object")
    obj.props["keys"] = NativeFunc(lambda this, a: JSArray(list(a[0].keys())) if isinstance(a[0], dict) else JSArray(), "keys")
    g["Object"] = obj
    g["Error"] = NativeFunc(lambda this, a: {"message": js_to_str(a[0]) if a else "", "name": "Error"}, "Error")
    g["Date"] = NativeFunc(lambda this, a: {"getTime": NativeFunc(lambda t, b: int(time.time() * 1000), "getTime")}, "Date")
    g["parseFloat"] = NativeFunc(lambda this, a: _parse_float(js_to_str(a[0])), "parseFloat")
    g["parseInt"] = NativeFunc(lambda this, a: _parse_int(js_to_str(a[0]), a[1] if len(a) > 1 else 10), "parseInt")
    g["isNaN"] = NativeFunc(lambda this, a: (lambda v: v != v)(to_num(a[0])), "isNaN")
    g["isFinite"] = NativeFunc(lambda this, a: (lambda v: v == v and v not in (INF, -INF))(to_num(a[0])), "isFinite")
    g["console"] = {"log": NativeFunc(lambda this, a: I.prints.append(" ".join(js_to_str(x) for x in a)), "log")}


def _parse_float(s):
    s = s.strip()
    j, n = 0, len(s)
    if j < n and s[j] in "+-":
        j += 1
    k = j
    while k < n and s[k].isdigit():
        k += 1
    if k < n and s[k] == ".":
        k += 1
        while k < n and s[k].isdigit():
            k += 1
    if k > j and k < n and s[k] in "eE":
        q = k + 1
        if q < n and s[q] in "+-":
            q += 1
        if q < n and s[q].isdigit():
            while q < n and s[q].isdigit():
                q += 1
            k = q
    if k == j:
        if s[j:j + 8] == "Infinity":
            return -INF if s.startswith("-") else INF
        return NAN
    try:
        return float(s[:k])
    except ValueError:
        return NAN


def _parse_int(s, radix):
    s = s.strip()
    radix = int(to_num(radix)) or 10
    j, n = 0, len(s)
    neg = False
    if j < n and s[j] in "+-":
        neg = s[j] == "-"
        j += 1
    if radix == 16 and s[j:j + 2].lower() == "0x":
        j += 2
    digs = "0123456789abcdefghijklmnopqrstuvwxyz"[:radix]
    k = j
    while k < n and s[k].lower() in digs:
        k += 1
    if k == j:
        return NAN
    v = int(s[j:k], radix)
    return -v if neg else v


def js_stack(tb):
    """qualified names of the JS functions active in a Python traceback"""
    out = []
    while tb is not None:
        fr = tb.tb_frame
        if fr.f_code.co_name == "call" and isinstance(fr.f_locals.get("self"), JSFunc):
            out.append(fr.f_locals["self"].qname)
        tb = tb.tb_next
    return out
