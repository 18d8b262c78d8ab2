"""minijs -- a small ECMAScript-5 interpreter, TEST INFRASTRUCTURE ONLY.

Why it exists: the reference (rasmusab/bayes.js) is JavaScript and this image has no JS engine (no node, d8, V8, QuickJS).
To pin the C oracle against the *real* reference rather than against a reading of it, this interpreter executes the
UNMODIFIED files /root/reference/mcmc.js, distributions.js and tests/test_data.js; oracle/minijs/make_golden.py drives it
with Math.random replaced by the Philox stream of DESIGN.md and writes tests/golden/*.json.

Coverage: the ES5 subset those files use -- closures, prototypes, `new`, `this`, call/apply, object/array literals,
for / for-in / do-while / while / switch / throw / try, typeof / instanceof / in / delete, ++/--, compound assignment,
comma and conditional expressions, loose and strict equality, string concatenation with array-to-string conversion.
Numbers are IEEE doubles (Python floats); Math.log / Math.exp are injected by the caller (the fdlibm restatement in
oracle/liboracle.so, i.e. what V8 computes), so values are comparable bit for bit with the C oracle.
"""
from __future__ import annotations

import math
import re
from typing import Any, Callable, Dict, List, Optional


# ------------------------------------------------------------------------------------------------------------------
# values
# ------------------------------------------------------------------------------------------------------------------
class Undefined:
    _inst = None

    def __new__(cls):
        if cls._inst is None:
            cls._inst = object.__new__(cls)
        return cls._inst

    def __repr__(self): return "undefined"
    def __bool__(self): return False


undefined = Undefined()


class JSObject:
    def __init__(self, proto: Optional["JSObject"] = None):
        self.props: Dict[str, Any] = {}
        self.proto = proto
        self.cls = "Object"

    def get(self, key: str):
        o = self
        while o is not None:
            if key in o.props:
                return o.props[key]
            o = o.proto
        return undefined

    def put(self, key: str, val):
        self.props[key] = val

    def has_own(self, key: str) -> bool:
        return key in self.props

    def has(self, key: str) -> bool:
        o = self
        while o is not None:
            if key in o.props:
                return True
            o = o.proto
        return False

    def delete(self, key: str):
        self.props.pop(key, None)
        return True

    def keys(self) -> List[str]:
        return list(self.props.keys())

    def enum_keys(self) -> List[str]:
        out, seen, o = [], set(), self
        while o is not None:
            for k in o.keys():
                if k not in seen and not (isinstance(o, JSFunction) and k == "prototype") and k != "constructor":
                    seen.add(k)
                    out.append(k)
            o = o.proto
        return out


class JSArray(JSObject):
    def __init__(self, items=None, proto=None):
        super().__init__(proto)
        self.items: List[Any] = list(items) if items is not None else []
        self.cls = "Array"

    @staticmethod
    def _idx(key: str):
        if key.isdigit():
            return int(key)
        return None

    def get(self, key: str):
        if key == "length":
            return float(len(self.items))
        i = self._idx(key)
        if i is not None:
            return self.items[i] if i < len(self.items) else undefined
        return super().get(key)

    def put(self, key: str, val):
        if key == "length":
            n = int(val)
            if n < len(self.items):
                del self.items[n:]
            else:
                self.items.extend([undefined] * (n - len(self.items)))
            return
        i = self._idx(key)
        if i is not None:
            if i >= len(self.items):
                self.items.extend([undefined] * (i + 1 - len(self.items)))
            self.items[i] = val
            return
        super().put(key, val)

    def has_own(self, key: str) -> bool:
        i = self._idx(key)
        if i is not None:
            return i < len(self.items) and self.items[i] is not undefined
        return key == "length" or super().has_own(key)

    def has(self, key: str) -> bool:
        return self.has_own(key) or super().has(key)

    def keys(self) -> List[str]:
        return [str(i) for i, v in enumerate(self.items) if v is not undefined] + list(self.props.keys())


class JSFunction(JSObject):
    def __init__(self, interp, params, body, env, name="", native: Optional[Callable] = None, proto=None):
        super().__init__(proto)
        self.interp, self.params, self.body, self.env, self.name, self.native = interp, params, body, env, name, native
        self.cls = "Function"

    def call(self, this, args: List[Any]):
        if self.native is not None:
            return self.native(this, args)
        return self.interp.call_function(self, this, args)


class JSThrow(Exception):
    def __init__(self, value):
        super().__init__(repr(value))
        self.value = value


class _Break(Exception):
    pass


class _Continue(Exception):
    pass


class _Return(Exception):
    def __init__(self, value):
        self.value = value


# ------------------------------------------------------------------------------------------------------------------
# lexer
# ------------------------------------------------------------------------------------------------------------------
_TOKEN = re.compile(r"""
    (?P<ws>\s+|//[^\n]*|/\*.*?\*/)
  | (?P<num>0[xX][0-9a-fA-F]+|(?:\d+\.?\d*(?:[eE][+-]?\d+)?|\.\d+(?:[eE][+-]?\d+)?))
  | (?P<id>[A-Za-z_$][A-Za-z0-9_$]*)
  | (?P<str>"(?:[^"\\]|\\.)*"|'(?:[^'\\]|\\.)*')
  | (?P<op>===|!==|>>>|<<=|>>=|==|!=|<=|>=|&&|\|\||\+\+|--|\+=|-=|\*=|/=|%=|<<|>>|[{}()\[\];,<>+\-*/%&|^!~?:=.])
""", re.X | re.S)

_KEYWORDS = {"var", "function", "return", "if", "else", "for", "while", "do", "break", "continue", "new", "delete", "typeof",
             "instanceof", "in", "this", "null", "true", "false", "throw", "try", "catch", "finally", "switch", "case",
             "default", "void"}


class TokenList(list):
    """tokens as (kind, value) pairs; `spans[i]` is the (start, end) of token i in `src` (Function.prototype.toString)"""
    src = ""
    spans: list = []


def tokenize(src: str):
    toks = TokenList()
    toks.src, toks.spans = src, []
    pos = 0
    while pos < len(src):
        m = _TOKEN.match(src, pos)
        if not m:
            raise SyntaxError(f"minijs: cannot tokenize at {pos}: {src[pos:pos + 30]!r}")
        pos = m.end()
        if m.lastgroup == "ws":
            continue
        toks.spans.append((m.start(), m.end()))
        kind, text = m.lastgroup, m.group(m.lastgroup)
        if kind == "num":
            toks.append(("num", float(int(text, 16)) if text[:2] in ("0x", "0X") else float(text)))
        elif kind == "id":
            toks.append(("kw" if text in _KEYWORDS else "id", text))
        elif kind == "str":
            body = text[1:-1]
            body = re.sub(r"\\(.)", lambda mm: {"n": "\n", "t": "\t", "r": "\r", "0": "\0"}.get(mm.group(1), mm.group(1)), body)
            toks.append(("str", body))
        else:
            toks.append(("op", text))
    toks.append(("eof", None))
    toks.spans.append((len(src), len(src)))
    return toks


# ------------------------------------------------------------------------------------------------------------------
# parser (AST = tuples)
# ------------------------------------------------------------------------------------------------------------------
_BINPREC = {"||": 1, "&&": 2, "|": 3, "^": 4, "&": 5, "==": 6, "!=": 6, "===": 6, "!==": 6, "<": 7, ">": 7, "<=": 7, ">=": 7,
            "instanceof": 7, "in": 7, "<<": 8, ">>": 8, ">>>": 8, "+": 9, "-": 9, "*": 10, "/": 10, "%": 10}


class Parser:
    def __init__(self, toks):
        self.t, self.i = toks, 0

    def _text(self, start_tok):
        """source text of tokens [start_tok, self.i)"""
        spans = getattr(self.t, "spans", None)
        if not spans:
            return None
        return self.t.src[spans[start_tok][0]:spans[self.i - 1][1]]

    def peek(self): return self.t[self.i]
    def next(self):
        tok = self.t[self.i]; self.i += 1
        return tok

    def at(self, kind, val=None):
        k, v = self.t[self.i]
        return k == kind and (val is None or v == val)

    def eat(self, kind, val=None):
        if self.at(kind, val):
            return self.next()
        return None

    def expect(self, kind, val=None):
        if not self.at(kind, val):
            raise SyntaxError(f"minijs: expected {val or kind}, got {self.peek()} at token {self.i}")
        return self.next()

    def program(self):
        body = []
        while not self.at("eof"):
            body.append(self.statement())
        return ("block", body)

    def semicolon(self):
        self.eat("op", ";")          # automatic semicolon insertion: the sources always terminate or sit before } / newline

    def statement(self):
        k, v = self.peek()
        if k == "op" and v == "{":
            return self.block()
        if k == "op" and v == ";":
            self.next()
            return ("empty",)
        if k == "kw":
            if v == "var":
                self.next()
                decls = self.var_decls()
                self.semicolon()
                return ("var", decls)
            if v == "function" and self.t[self.i + 1][0] == "id":
                self.next()
                name = self.next()[1]
                start = self.i - 2
                params, body = self.func_rest()
                return ("funcdecl", name, params, body, self._text(start))
            if v == "return":
                self.next()
                e = None
                if not (self.at("op", ";") or self.at("op", "}") or self.at("eof")):
                    e = self.expression()
                self.semicolon()
                return ("return", e)
            if v == "if":
                self.next(); self.expect("op", "(")
                c = self.expression(); self.expect("op", ")")
                a = self.statement()
                b = self.statement() if self.eat("kw", "else") else None
                return ("if", c, a, b)
            if v == "for":
                return self.for_stmt()
            if v == "while":
                self.next(); self.expect("op", "(")
                c = self.expression(); self.expect("op", ")")
                return ("while", c, self.statement())
            if v == "do":
                self.next()
                body = self.statement()
                self.expect("kw", "while"); self.expect("op", "(")
                c = self.expression(); self.expect("op", ")")
                self.semicolon()
                return ("dowhile", body, c)
            if v == "break":
                self.next(); self.semicolon()
                return ("break",)
            if v == "continue":
                self.next(); self.semicolon()
                return ("continue",)
            if v == "throw":
                self.next()
                e = self.expression(); self.semicolon()
                return ("throw", e)
            if v == "try":
                self.next()
                blk = self.block()
                param = handler = final = None
                if self.eat("kw", "catch"):
                    self.expect("op", "("); param = self.expect("id")[1]; self.expect("op", ")")
                    handler = self.block()
                if self.eat("kw", "finally"):
                    final = self.block()
                return ("try", blk, param, handler, final)
            if v == "switch":
                self.next(); self.expect("op", "(")
                disc = self.expression(); self.expect("op", ")"); self.expect("op", "{")
                cases = []
                while not self.eat("op", "}"):
                    if self.eat("kw", "default"):
                        test = None
                    else:
                        self.expect("kw", "case"); test = self.expression()
                    self.expect("op", ":")
                    body = []
                    while not (self.at("kw", "case") or self.at("kw", "default") or self.at("op", "}")):
                        body.append(self.statement())
                    cases.append((test, body))
                return ("switch", disc, cases)
        e = self.expression()
        self.semicolon()
        return ("expr", e)

    def block(self):
        self.expect("op", "{")
        body = []
        while not self.eat("op", "}"):
            body.append(self.statement())
        return ("block", body)

    def var_decls(self, no_in=False):
        decls = []
        while True:
            name = self.expect("id")[1]
            init = self.assignment(no_in) if self.eat("op", "=") else None
            decls.append((name, init))
            if not self.eat("op", ","):
                return decls

    def for_stmt(self):
        self.next(); self.expect("op", "(")
        init = None
        if self.eat("kw", "var"):
            decls = self.var_decls(no_in=True)
            if self.eat("kw", "in"):
                obj = self.expression(); self.expect("op", ")")
                return ("forin", ("var", decls), decls[0][0], obj, self.statement())
            init = ("var", decls)
        elif not self.at("op", ";"):
            e = self.expression(no_in=True)
            if self.eat("kw", "in"):
                obj = self.expression(); self.expect("op", ")")
                return ("forin", None, e, obj, self.statement())
            init = ("expr", e)
        self.expect("op", ";")
        test = None if self.at("op", ";") else self.expression()
        self.expect("op", ";")
        update = None if self.at("op", ")") else self.expression()
        self.expect("op", ")")
        return ("for", init, test, update, self.statement())

    def func_rest(self):
        self.expect("op", "(")
        params = []
        while not self.eat("op", ")"):
            params.append(self.expect("id")[1])
            self.eat("op", ",")
        body = self.block()
        return params, body

    # -- expressions ----------------------------------------------------------------------------------------------
    def expression(self, no_in=False):
        e = self.assignment(no_in)
        while self.eat("op", ","):
            e = ("comma", e, self.assignment(no_in))
        return e

    def assignment(self, no_in=False):
        left = self.conditional(no_in)
        k, v = self.peek()
        if k == "op" and v in ("=", "+=", "-=", "*=", "/=", "%="):
            self.next()
            right = self.assignment(no_in)
            return ("assign", v, left, right)
        return left

    def conditional(self, no_in=False):
        c = self.binary(0, no_in)
        if self.eat("op", "?"):
            a = self.assignment()
            self.expect("op", ":")
            b = self.assignment(no_in)
            return ("cond", c, a, b)
        return c

    def binary(self, minprec, no_in=False):
        left = self.unary()
        while True:
            k, v = self.peek()
            if not ((k == "op" and v in _BINPREC) or (k == "kw" and v in ("instanceof", "in"))):
                return left
            if v == "in" and no_in:
                return left
            prec = _BINPREC[v]
            if prec <= minprec:
                return left
            self.next()
            right = self.binary(prec, no_in)
            left = ("logical" if v in ("&&", "||") else "bin", v, left, right)

    def unary(self):
        k, v = self.peek()
        if k == "op" and v in ("!", "-", "+", "~"):
            self.next()
            return ("unary", v, self.unary())
        if k == "op" and v in ("++", "--"):
            self.next()
            return ("update", v, True, self.unary())
        if k == "kw" and v in ("typeof", "delete", "void"):
            self.next()
            return ("unary", v, self.unary())
        return self.postfix()

    def postfix(self):
        e = self.call_member()
        k, v = self.peek()
        if k == "op" and v in ("++", "--"):
            self.next()
            return ("update", v, False, e)
        return e

    def call_member(self):
        if self.at("kw", "new"):
            self.next()
            callee = self.member_only()
            args = self.arguments() if self.at("op", "(") else []
            e = ("new", callee, args)
        else:
            e = self.primary()
        while True:
            if self.eat("op", "."):
                name = self.next()[1]
                e = ("member", e, ("str", name))
            elif self.eat("op", "["):
                idx = self.expression()
                self.expect("op", "]")
                e = ("member", e, idx)
            elif self.at("op", "("):
                e = ("call", e, self.arguments())
            else:
                return e

    def member_only(self):
        e = self.primary()
        while True:
            if self.eat("op", "."):
                e = ("member", e, ("str", self.next()[1]))
            elif self.eat("op", "["):
                idx = self.expression(); self.expect("op", "]")
                e = ("member", e, idx)
            else:
                return e

    def arguments(self):
        self.expect("op", "(")
        args = []
        while not self.eat("op", ")"):
            args.append(self.assignment())
            self.eat("op", ",")
        return args

    def primary(self):
        k, v = self.next()
        if k == "num": return ("num", v)
        if k == "str": return ("str", v)
        if k == "id": return ("ident", v)
        if k == "kw":
            if v == "this": return ("this",)
            if v == "null": return ("null",)
            if v == "true": return ("bool", True)
            if v == "false": return ("bool", False)
            if v == "function":
                start = self.i - 1
                name = self.next()[1] if self.at("id") else ""
                params, body = self.func_rest()
                return ("func", name, params, body, self._text(start))
        if k == "op":
            if v == "(":
                e = self.expression()
                self.expect("op", ")")
                return e
            if v == "[":
                items = []
                while not self.eat("op", "]"):
                    items.append(self.assignment())
                    self.eat("op", ",")
                return ("array", items)
            if v == "{":
                props = []
                while not self.eat("op", "}"):
                    kk, kv = self.next()
                    key = _num_key(kv) if kk == "num" else str(kv)
                    self.expect("op", ":")
                    props.append((key, self.assignment()))
                    self.eat("op", ",")
                return ("object", props)
        raise SyntaxError(f"minijs: unexpected token {(k, v)} at {self.i}")


def _num_key(v: float) -> str:
    return str(int(v)) if v == int(v) and abs(v) < 1e21 else repr(v)


# ------------------------------------------------------------------------------------------------------------------
# interpreter
# ------------------------------------------------------------------------------------------------------------------
class Env:
    __slots__ = ("vars", "parent")

    def __init__(self, parent=None):
        self.vars: Dict[str, Any] = {}
        self.parent = parent

    def lookup(self, name):
        e = self
        while e is not None:
            if name in e.vars:
                return e
            e = e.parent
        return None


class Interpreter:
    def __init__(self, math_log=math.log, math_exp=math.exp, random: Callable[[], float] = None):
        self.object_proto = JSObject(None)
        self.function_proto = JSObject(self.object_proto)
        self.array_proto = JSObject(self.object_proto)
        self.global_env = Env()
        self.global_obj = JSObject(self.object_proto)
        self.math_log, self.math_exp = math_log, math_exp
        self.random = random or (lambda: 0.5)
        self._setup_globals()

    # -- conversions ----------------------------------------------------------------------------------------------
    def to_bool(self, v) -> bool:
        if v is undefined or v is None: return False
        if isinstance(v, bool): return v
        if isinstance(v, float): return not (v == 0 or v != v)
        if isinstance(v, str): return v != ""
        return True

    def to_num(self, v) -> float:
        if isinstance(v, bool): return 1.0 if v else 0.0
        if isinstance(v, float): return v
        if v is undefined: return math.nan
        if v is None: return 0.0
        if isinstance(v, str):
            s = v.strip()
            if s == "": return 0.0
            try:
                return float(int(s, 16)) if s[:2] in ("0x", "0X") else float(s)
            except ValueError:
                return math.nan
        if isinstance(v, JSArray):
            return self.to_num(self.to_str(v))
        return math.nan

    def num_str(self, v: float) -> str:
        if v != v: return "NaN"
        if v == math.inf: return "Infinity"
        if v == -math.inf: return "-Infinity"
        if v == int(v) and abs(v) < 1e21: return str(int(v))
        return repr(v)

    def to_str(self, v) -> str:
        if isinstance(v, str): return v
        if isinstance(v, bool): return "true" if v else "false"
        if isinstance(v, float): return self.num_str(v)
        if v is undefined: return "undefined"
        if v is None: return "null"
        if isinstance(v, JSArray):
            return ",".join("" if (x is undefined or x is None) else self.to_str(x) for x in v.items)
        if isinstance(v, JSFunction): return getattr(v, "source", None) or ("function " + v.name + "() { [native code] }")
        return "[object Object]"

    def to_key(self, v) -> str:
        if isinstance(v, float):
            return _num_key(v)
        return self.to_str(v)

    def typeof(self, v) -> str:
        if v is undefined: return "undefined"
        if v is None: return "object"
        if isinstance(v, bool): return "boolean"
        if isinstance(v, float): return "number"
        if isinstance(v, str): return "string"
        if isinstance(v, JSFunction): return "function"
        return "object"

    def loose_eq(self, a, b) -> bool:
        if (a is undefined or a is None) and (b is undefined or b is None): return True
        if a is undefined or a is None or b is undefined or b is None: return False
        if isinstance(a, JSObject) and isinstance(b, JSObject): return a is b
        if isinstance(a, JSObject): a = self.to_str(a)
        if isinstance(b, JSObject): b = self.to_str(b)
        if isinstance(a, str) and isinstance(b, str): return a == b
        return self.to_num(a) == self.to_num(b)

    def strict_eq(self, a, b) -> bool:
        if isinstance(a, bool) or isinstance(b, bool):
            return isinstance(a, bool) and isinstance(b, bool) and a == b
        if isinstance(a, float) and isinstance(b, float): return a == b
        if isinstance(a, str) and isinstance(b, str): return a == b
        return a is b

    # -- property access --------------------------------------------------------------------------------------------
    def get_prop(self, obj, key: str):
        if isinstance(obj, JSObject):
            return obj.get(key)
        if isinstance(obj, str):
            if key == "length": return float(len(obj))
            if key.isdigit(): return obj[int(key)] if int(key) < len(obj) else undefined
            return self.string_proto.get(key)
        if isinstance(obj, float):
            return self.number_proto.get(key)
        if obj is undefined or obj is None:
            raise JSThrow(f"TypeError: Cannot read property '{key}' of {self.to_str(obj)}")
        return undefined

    def make_native(self, name, fn) -> JSFunction:
        return JSFunction(self, [], None, None, name=name, native=fn, proto=self.function_proto)

    def new_array(self, items) -> JSArray:
        return JSArray(items, self.array_proto)

    def new_object(self) -> JSObject:
        return JSObject(self.object_proto)

    # -- globals ----------------------------------------------------------------------------------------------------
    def _setup_globals(self):
        g = self.global_env.vars
        nat = self.make_native
        op, fp, ap = self.object_proto, self.function_proto, self.array_proto
        self.string_proto = JSObject(op)
        self.number_proto = JSObject(op)
        op.put("hasOwnProperty", nat("hasOwnProperty", lambda this, a: isinstance(this, JSObject) and this.has_own(self.to_key(a[0]))))

        def obj_to_string(this, a):
            if isinstance(this, JSArray): return "[object Array]"
            if isinstance(this, JSFunction): return "[object Function]"
            if this is undefined: return "[object Undefined]"
            if this is None: return "[object Null]"
            if isinstance(this, float): return "[object Number]"
            if isinstance(this, str): return "[object String]"
            return "[object " + getattr(this, "cls", "Object") + "]"
        op.put("toString", nat("toString", obj_to_string))
        fp.put("call", nat("call", lambda this, a: this.call(a[0] if a else undefined, a[1:])))
        fp.put("apply", nat("apply", lambda this, a: this.call(a[0] if a else undefined, list(a[1].items) if len(a) > 1 and isinstance(a[1], JSArray) else [])))

        def arr_push(this, a):
            this.items.extend(a)
            return float(len(this.items))

        def arr_slice(this, a):
            n = len(this.items)
            s = int(self.to_num(a[0])) if a and a[0] is not undefined else 0
            e = int(self.to_num(a[1])) if len(a) > 1 and a[1] is not undefined else n
            if s < 0: s += n
            if e < 0: e += n
            return self.new_array(this.items[s:e])

        def arr_concat(this, a):
            out = list(this.items)
            for x in a:
                out.extend(x.items) if isinstance(x, JSArray) else out.append(x)
            return self.new_array(out)

        def arr_join(this, a):
            sep = self.to_str(a[0]) if a and a[0] is not undefined else ","
            return sep.join("" if (x is undefined or x is None) else self.to_str(x) for x in this.items)

        def arr_index_of(this, a):
            for i, x in enumerate(this.items):
                if self.strict_eq(x, a[0]): return float(i)
            return -1.0
        ap.put("push", nat("push", arr_push)); ap.put("slice", nat("slice", arr_slice)); ap.put("concat", nat("concat", arr_concat))
        ap.put("join", nat("join", arr_join)); ap.put("indexOf", nat("indexOf", arr_index_of))
        ap.put("toString", nat("toString", lambda this, a: self.to_str(this)))

        def make_ctor(name, fn, proto):
            f = nat(name, fn)
            f.put("prototype", proto)
            proto.put("constructor", f)
            g[name] = f
            return f

        def object_ctor(this, a):
            return self.new_object()
        Object = make_ctor("Object", object_ctor, op)
        Object.put("keys", nat("keys", lambda this, a: self.new_array([k for k in a[0].keys() if not (isinstance(a[0], JSArray) and k == "length")])))
        Object.put("create", nat("create", lambda this, a: JSObject(a[0] if isinstance(a[0], JSObject) else None)))

        def array_ctor(this, a):
            if len(a) == 1 and isinstance(a[0], float):
                return self.new_array([undefined] * int(a[0]))
            return self.new_array(a)
        Array = make_ctor("Array", array_ctor, ap)
        Array.put("isArray", nat("isArray", lambda this, a: isinstance(a[0], JSArray) if a else False))
        def function_ctor(this, a):
            """new Function(p1, ..., pn, body): the body runs in the GLOBAL scope, like the real constructor"""
            parts = [self.to_str(x) for x in a]
            params, body = parts[:-1], (parts[-1] if parts else "")
            src = "(function anonymous(" + ",".join(params) + "\n) {\n" + body + "\n})"
            ast = Parser(tokenize(src)).expression()
            return self.eval(ast, self.global_env)
        make_ctor("Function", function_ctor, fp)
        fp.put("toString", nat("toString", lambda this, a: self.to_str(this)))
        sp = self.string_proto

        def _s(this): return this if isinstance(this, str) else self.to_str(this)

        def _i(a, k, default):
            if len(a) <= k or a[k] is undefined: return default
            v = self.to_num(a[k])
            return default if v != v else int(v)
        sp.put("charAt", nat("charAt", lambda this, a: (_s(this)[_i(a, 0, 0)] if 0 <= _i(a, 0, 0) < len(_s(this)) else "")))
        sp.put("charCodeAt", nat("charCodeAt", lambda this, a: (float(ord(_s(this)[_i(a, 0, 0)])) if 0 <= _i(a, 0, 0) < len(_s(this)) else math.nan)))

        def str_substring(this, a):
            st = _s(this)
            b, e = max(0, min(len(st), _i(a, 0, 0))), max(0, min(len(st), _i(a, 1, len(st))))
            if b > e: b, e = e, b
            return st[b:e]

        def str_slice(this, a):
            st = _s(this)
            n = len(st)
            b, e = _i(a, 0, 0), _i(a, 1, n)
            if b < 0: b = max(0, n + b)
            if e < 0: e = max(0, n + e)
            return st[min(b, n):min(e, n)] if b < e else ""
        sp.put("substring", nat("substring", str_substring))
        sp.put("slice", nat("slice", str_slice))
        sp.put("indexOf", nat("indexOf", lambda this, a: float(_s(this).find(self.to_str(a[0]) if a else "undefined", _i(a, 1, 0)))))
        sp.put("toString", nat("toString", lambda this, a: _s(this)))
        sp.put("split", nat("split", lambda this, a: self.new_array(list(_s(this)) if (a and a[0] == "") else _s(this).split(self.to_str(a[0])) if a else [_s(this)])))
        sp.put("replace", nat("replace", lambda this, a: _s(this).replace(self.to_str(a[0]), self.to_str(a[1]), 1)))
        sp.put("toLowerCase", nat("toLowerCase", lambda this, a: _s(this).lower()))
        self.number_proto.put("toString", nat("toString", lambda this, a: self.to_str(this)))
        self.number_proto.put("toFixed", nat("toFixed", lambda this, a: format(self.to_num(this), "." + str(int(self.to_num(a[0])) if a else 0) + "f")))

        def arr_pop(this, a):
            return this.items.pop() if this.items else undefined

        def arr_shift(this, a):
            return this.items.pop(0) if this.items else undefined

        def arr_sort(this, a):
            import functools
            if a and isinstance(a[0], JSFunction):
                this.items.sort(key=functools.cmp_to_key(lambda x, y: (lambda r: -1 if r < 0 else (1 if r > 0 else 0))(self.to_num(a[0].call(undefined, [x, y])))))
            else:
                this.items.sort(key=lambda x: self.to_str(x))
            return this

        def arr_reverse(this, a):
            this.items.reverse()
            return this
        def arr_splice(this, a):
            n = len(this.items)
            start = int(self.to_num(a[0])) if a else 0
            if start < 0: start = max(0, n + start)
            start = min(start, n)
            count = n - start if len(a) < 2 else max(0, min(int(self.to_num(a[1])), n - start))
            removed = this.items[start:start + count]
            this.items[start:start + count] = list(a[2:])
            return self.new_array(removed)
        ap.put("splice", nat("splice", arr_splice))
        ap.put("pop", nat("pop", arr_pop)); ap.put("shift", nat("shift", arr_shift)); ap.put("sort", nat("sort", arr_sort)); ap.put("reverse", nat("reverse", arr_reverse))
        make_ctor("Number", lambda this, a: self.to_num(a[0]) if a else 0.0, self.number_proto)
        make_ctor("String", lambda this, a: self.to_str(a[0]) if a else "", self.string_proto)
        make_ctor("Date", lambda this, a: self.new_object(), JSObject(op))
        make_ctor("RegExp", lambda this, a: self.new_object(), JSObject(op))
        make_ctor("Error", lambda this, a: self.new_object(), JSObject(op))

        M = self.new_object()

        def js_round(x):
            if x != x or x in (math.inf, -math.inf): return x
            r = float(math.ceil(x))
            return r - 1.0 if r - 0.5 > x else r

        def js_pow(x, y):
            if y == 2.0: return x * x
            try:
                return math.pow(x, y)
            except (OverflowError, ValueError):
                return math.inf if x > 0 else math.nan

        def js_minmax(args, pick):
            vals = [self.to_num(v) for v in args]
            if any(v != v for v in vals): return math.nan
            return pick(vals) if vals else (-math.inf if pick is max else math.inf)
        M.put("PI", math.pi); M.put("E", math.e)
        M.put("random", nat("random", lambda this, a: self.random()))
        M.put("floor", nat("floor", lambda this, a: float(math.floor(self.to_num(a[0]))) if math.isfinite(self.to_num(a[0])) else self.to_num(a[0])))
        M.put("ceil", nat("ceil", lambda this, a: float(math.ceil(self.to_num(a[0]))) if math.isfinite(self.to_num(a[0])) else self.to_num(a[0])))
        M.put("round", nat("round", lambda this, a: js_round(self.to_num(a[0]))))
        M.put("log", nat("log", lambda this, a: self.math_log(self.to_num(a[0]))))
        M.put("exp", nat("exp", lambda this, a: self.math_exp(self.to_num(a[0]))))
        M.put("sqrt", nat("sqrt", lambda this, a: math.sqrt(self.to_num(a[0])) if self.to_num(a[0]) >= 0 else math.nan))
        M.put("abs", nat("abs", lambda this, a: abs(self.to_num(a[0]))))
        M.put("pow", nat("pow", lambda this, a: js_pow(self.to_num(a[0]), self.to_num(a[1]))))
        M.put("max", nat("max", lambda this, a: js_minmax(a, max)))
        M.put("min", nat("min", lambda this, a: js_minmax(a, min)))
        g["Math"] = M
        g["Infinity"] = math.inf
        g["NaN"] = math.nan
        g["undefined"] = undefined
        g["isNaN"] = nat("isNaN", lambda this, a: self.to_num(a[0]) != self.to_num(a[0]))
        g["parseFloat"] = nat("parseFloat", lambda this, a: self.to_num(a[0]))
        g["isFinite"] = nat("isFinite", lambda this, a: math.isfinite(self.to_num(a[0])))

    # -- running ----------------------------------------------------------------------------------------------------
    def run(self, src: str, this=None):
        ast = Parser(tokenize(src)).program()
        self._hoist(ast[1], self.global_env)
        if not hasattr(self, "_this_stack"):
            self._this_stack = []
        self._this_stack.append(this if this is not None else self.global_obj)      # run() may nest (a native `require` loads a module)
        try:
            self.exec_block(ast[1], self.global_env)
        except _Return:
            pass
        finally:
            self._this_stack.pop()

    def get_global(self, name):
        return self.global_env.vars.get(name, undefined)

    def set_global(self, name, v):
        self.global_env.vars[name] = v

    def _hoist(self, stmts, env: Env):
        """var and function declarations are hoisted to the enclosing function scope."""
        for s in stmts:
            self._hoist_stmt(s, env)

    def _hoist_stmt(self, s, env):
        if s is None: return
        k = s[0]
        if k == "var":
            for name, _ in s[1]:
                env.vars.setdefault(name, undefined)
        elif k == "funcdecl":
            env.vars[s[1]] = self.make_function(s[1], s[2], s[3], env, s[4] if len(s) > 4 else None)
        elif k == "block":
            self._hoist(s[1], env)
        elif k == "if":
            self._hoist_stmt(s[2], env); self._hoist_stmt(s[3], env)
        elif k == "for":
            self._hoist_stmt(s[1], env); self._hoist_stmt(s[4], env)
        elif k == "forin":
            self._hoist_stmt(s[1], env); self._hoist_stmt(s[4], env)
        elif k in ("while",):
            self._hoist_stmt(s[2], env)
        elif k == "dowhile":
            self._hoist_stmt(s[1], env)
        elif k == "try":
            self._hoist_stmt(s[1], env); self._hoist_stmt(s[3], env); self._hoist_stmt(s[4], env)
        elif k == "switch":
            for _, body in s[2]:
                self._hoist(body, env)

    def make_function(self, name, params, body, env, source=None) -> JSFunction:
        f = JSFunction(self, params, body, env, name=name, proto=self.function_proto)
        f.source = source
        proto = JSObject(self.object_proto)
        proto.put("constructor", f)
        f.put("prototype", proto)
        return f

    def call_function(self, f: JSFunction, this, args):
        env = Env(f.env)
        for i, p in enumerate(f.params):
            env.vars[p] = args[i] if i < len(args) else undefined
        env.vars["arguments"] = self.new_array(args)
        if f.name and f.name not in env.vars:
            env.vars[f.name] = f
        self._hoist(f.body[1], env)
        self._this_stack.append(this)
        try:
            self.exec_block(f.body[1], env)
        except _Return as r:
            return r.value
        finally:
            self._this_stack.pop()
        return undefined

    def exec_block(self, stmts, env):
        for s in stmts:
            self.exec(s, env)

    def exec(self, s, env):
        k = s[0]
        if k == "expr":
            self.eval(s[1], env)
        elif k == "var":
            for name, init in s[1]:
                if init is not None:
                    env.lookup(name).vars[name] = self.eval(init, env)
        elif k == "return":
            raise _Return(self.eval(s[1], env) if s[1] is not None else undefined)
        elif k == "if":
            if self.to_bool(self.eval(s[1], env)):
                self.exec(s[2], env)
            elif s[3] is not None:
                self.exec(s[3], env)
        elif k == "block":
            self.exec_block(s[1], env)
        elif k == "for":
            if s[1] is not None: self.exec(s[1], env)
            while s[2] is None or self.to_bool(self.eval(s[2], env)):
                try:
                    self.exec(s[4], env)
                except _Break:
                    break
                except _Continue:
                    pass
                if s[3] is not None: self.eval(s[3], env)
        elif k == "forin":
            obj = self.eval(s[3], env)
            keys = obj.enum_keys() if isinstance(obj, JSObject) else []
            if isinstance(obj, JSArray):
                keys = [kk for kk in keys if kk != "length"]
            for key in keys:
                if isinstance(obj, JSObject) and not obj.has(key):
                    continue
                if isinstance(s[2], str):
                    (env.lookup(s[2]) or self.global_env).vars[s[2]] = key
                else:
                    self.assign_to(s[2], key, env)
                try:
                    self.exec(s[4], env)
                except _Break:
                    break
                except _Continue:
                    continue
        elif k == "while":
            while self.to_bool(self.eval(s[1], env)):
                try:
                    self.exec(s[2], env)
                except _Break:
                    break
                except _Continue:
                    continue
        elif k == "dowhile":
            while True:
                try:
                    self.exec(s[1], env)
                except _Break:
                    break
                except _Continue:
                    pass
                if not self.to_bool(self.eval(s[2], env)):
                    break
        elif k == "break":
            raise _Break()
        elif k == "continue":
            raise _Continue()
        elif k == "throw":
            raise JSThrow(self.eval(s[1], env))
        elif k == "try":
            try:
                try:
                    self.exec(s[1], env)
                except JSThrow as e:
                    if s[3] is None:
                        raise
                    cenv = Env(env)
                    cenv.vars[s[2]] = e.value
                    self.exec(s[3], cenv)
            finally:
                if s[4] is not None:
                    self.exec(s[4], env)
        elif k == "switch":
            d = self.eval(s[1], env)
            matched = False
            try:
                for test, body in s[2]:
                    if not matched and test is not None and self.strict_eq(d, self.eval(test, env)):
                        matched = True
                    if matched:
                        self.exec_block(body, env)
                if not matched:
                    run = False
                    for test, body in s[2]:
                        if test is None: run = True
                        if run: self.exec_block(body, env)
            except _Break:
                pass
        elif k in ("funcdecl", "empty"):
            pass
        else:
            raise NotImplementedError(k)

    # -- expressions ----------------------------------------------------------------------------------------------------
    def eval(self, e, env):
        k = e[0]
        if k == "num" or k == "str" or k == "bool": return e[1]
        if k == "null": return None
        if k == "ident":
            scope = env.lookup(e[1])
            if scope is None:
                raise JSThrow(f"ReferenceError: {e[1]} is not defined")
            return scope.vars[e[1]]
        if k == "this": return self._this_stack[-1]
        if k == "member":
            obj = self.eval(e[1], env)
            return self.get_prop(obj, self.to_key(self.eval(e[2], env)))
        if k == "call":
            callee = e[1]
            if callee[0] == "member":
                this = self.eval(callee[1], env)
                f = self.get_prop(this, self.to_key(self.eval(callee[2], env)))
            else:
                this = undefined
                f = self.eval(callee, env)
            args = [self.eval(a, env) for a in e[2]]
            if not isinstance(f, JSFunction):
                raise JSThrow("TypeError: " + self.to_str(callee[2][1] if callee[0] == "member" and callee[2][0] == "str" else "value") + " is not a function")
            return f.call(this, args)
        if k == "new":
            f = self.eval(e[1], env)
            args = [self.eval(a, env) for a in e[2]]
            if not isinstance(f, JSFunction):
                raise JSThrow("TypeError: not a constructor")
            if f.native is not None:
                return f.native(undefined, args)
            proto = f.get("prototype")
            obj = JSObject(proto if isinstance(proto, JSObject) else self.object_proto)
            r = f.call(obj, args)
            return r if isinstance(r, JSObject) else obj
        if k == "func":
            return self.make_function(e[1], e[2], e[3], env, e[4] if len(e) > 4 else None)
        if k == "array":
            return self.new_array([self.eval(x, env) for x in e[1]])
        if k == "object":
            o = self.new_object()
            for key, ve in e[1]:
                o.put(key, self.eval(ve, env))
            return o
        if k == "bin": return self.binop(e[1], self.eval(e[2], env), self.eval(e[3], env))
        if k == "logical":
            left = self.eval(e[2], env)
            if e[1] == "&&":
                return self.eval(e[3], env) if self.to_bool(left) else left
            return left if self.to_bool(left) else self.eval(e[3], env)
        if k == "unary":
            op = e[1]
            if op == "typeof":
                if e[2][0] == "ident" and env.lookup(e[2][1]) is None:
                    return "undefined"
                return self.typeof(self.eval(e[2], env))
            if op == "delete":
                t = e[2]
                if t[0] == "member":
                    obj = self.eval(t[1], env)
                    if isinstance(obj, JSObject):
                        return obj.delete(self.to_key(self.eval(t[2], env)))
                return True
            v = self.eval(e[2], env)
            if op == "!": return not self.to_bool(v)
            if op == "-": return -self.to_num(v)
            if op == "+": return self.to_num(v)
            if op == "void": return undefined
            if op == "~": return float(~int(self.to_num(v)))
        if k == "update":
            old = self.to_num(self.eval(e[3], env))
            new = old + 1 if e[1] == "++" else old - 1
            self.assign_to(e[3], new, env)
            return new if e[2] else old
        if k == "assign":
            op = e[1]
            if op == "=":
                v = self.eval(e[3], env)
            else:
                v = self.binop(op[:-1], self.eval(e[2], env), self.eval(e[3], env))
            self.assign_to(e[2], v, env)
            return v
        if k == "cond":
            return self.eval(e[2], env) if self.to_bool(self.eval(e[1], env)) else self.eval(e[3], env)
        if k == "comma":
            self.eval(e[1], env)
            return self.eval(e[2], env)
        raise NotImplementedError(k)

    def assign_to(self, target, v, env):
        if target[0] == "ident":
            scope = env.lookup(target[1]) or self.global_env      # sloppy-mode implicit global (tests/test_data.js:98 `x1 = ...`)
            scope.vars[target[1]] = v
        elif target[0] == "member":
            obj = self.eval(target[1], env)
            if not isinstance(obj, JSObject):
                raise JSThrow("TypeError: cannot set property of " + self.to_str(obj))
            obj.put(self.to_key(self.eval(target[2], env)), v)
        else:
            raise JSThrow("ReferenceError: invalid assignment target")

    def binop(self, op, a, b):
        if op == "+":
            if isinstance(a, JSObject): a = self.to_str(a)
            if isinstance(b, JSObject): b = self.to_str(b)
            if isinstance(a, str) or isinstance(b, str):
                return self.to_str(a) + self.to_str(b)
            return self.to_num(a) + self.to_num(b)
        if op == "-": return self.to_num(a) - self.to_num(b)
        if op == "*": return self.to_num(a) * self.to_num(b)
        if op == "/":
            x, y = self.to_num(a), self.to_num(b)
            if y == 0:
                if x == 0 or x != x: return math.nan
                return math.copysign(math.inf, x) * math.copysign(1.0, y)
            return x / y
        if op == "%":
            x, y = self.to_num(a), self.to_num(b)
            if y == 0 or x != x or y != y or math.isinf(x): return math.nan
            return math.fmod(x, y)
        if op in ("<", ">", "<=", ">="):
            if isinstance(a, str) and isinstance(b, str):
                x, y = a, b
            else:
                x, y = self.to_num(a), self.to_num(b)
                if x != x or y != y: return False
            return {"<": x < y, ">": x > y, "<=": x <= y, ">=": x >= y}[op]
        if op == "==": return self.loose_eq(a, b)
        if op == "!=": return not self.loose_eq(a, b)
        if op == "===": return self.strict_eq(a, b)
        if op == "!==": return not self.strict_eq(a, b)
        if op == "instanceof":
            if not isinstance(b, JSFunction): raise JSThrow("TypeError: right-hand side of instanceof is not callable")
            proto = b.get("prototype")
            o = a.proto if isinstance(a, JSObject) else None
            while o is not None:
                if o is proto: return True
                o = o.proto
            return False
        if op == "in":
            return isinstance(b, JSObject) and b.has(self.to_key(a))
        if op in ("&", "|", "^", "<<", ">>", ">>>"):
            x, y = int(self.to_num(a)) & 0xffffffff, int(self.to_num(b)) & 0xffffffff
            r = {"&": x & y, "|": x | y, "^": x ^ y, "<<": (x << (y & 31)) & 0xffffffff, ">>": x >> (y & 31), ">>>": x >> (y & 31)}[op]
            return float(r - (1 << 32) if (op != ">>>" and r & 0x80000000) else r)
        raise NotImplementedError(op)


# ------------------------------------------------------------------------------------------------------------------
# Python <-> JS conversion helpers
# ------------------------------------------------------------------------------------------------------------------
def to_py(v):
    if isinstance(v, JSArray): return [to_py(x) for x in v.items]
    if isinstance(v, JSFunction): return "<function>"
    if isinstance(v, JSObject): return {k: to_py(x) for k, x in v.props.items()}
    if v is undefined: return None
    return v


def to_js(interp: Interpreter, v):
    if isinstance(v, dict):
        o = interp.new_object()
        for k, x in v.items():
            o.put(str(k), to_js(interp, x))
        return o
    if isinstance(v, (list, tuple)):
        return interp.new_array([to_js(interp, x) for x in v])
    if isinstance(v, bool): return v
    if isinstance(v, (int, float)): return float(v)
    return v
