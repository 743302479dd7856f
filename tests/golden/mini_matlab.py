"""A small interpreter for the MATLAB subset the reference's hot-path functions are written in.

TEST INFRASTRUCTURE.  MATLAB / Octave are not installed in the build container, so the reference's own
``helperMIMOChannelEstimate.m`` and ``LMMSE_ce.m`` cannot be executed natively.  This module executes their
UNMODIFIED source text (read from /root/reference by tests/golden/make_golden.py) so that golden vectors can be
generated from the reference itself rather than from a hand translation:

* values are numpy arrays with MATLAB semantics: everything is at least 2-D (scalars are 1x1), column-major
  reshape / linear indexing, 1-based indices, implicit expansion, ``*`` is a matrix product unless one side is a
  scalar, ``'`` is the conjugate transpose and ``.'`` the plain transpose;
* statements: assignment (also indexed, ``A(:,j,i) = ...``, and multi-output ``[~, a, b] = f(...)``), ``for``,
  ``if / else``, several statements per line, ``...`` continuation, ``%`` comments, ``function`` files with
  sub-functions;
* expressions: numbers, ``i``/``j`` imaginary unit when not shadowed, matrix literals with ``,`` / blank / ``;``,
  ranges, ``end`` inside indices, struct fields, the operators ``+ - * / ^ .* ./ .^ ' .' : < > <= >= == ~= && || ~``;
* builtins: size zeros ones eye complex numel length squeeze reshape repmat inv sum mean norm sqrt conj transpose
  abs real imag floor mod isempty and a few more; char literals ('all') as option arguments; anything else (e.g. ``helperGetP``, a MathWorks example helper that is not
  in the reference repo) must be supplied by the caller as a Python callable.

It is deliberately general (a tokenizer + recursive-descent parser + tree-walking evaluator), not a line-by-line
special case, so that what runs is the reference's text.
"""
import re

import numpy as np

__all__ = ["MatlabFile", "MatlabError", "mat"]


class MatlabError(RuntimeError):
    pass


def mat(x):
    """numpy value with MATLAB shape rules (>= 2-D)."""
    a = np.asarray(x)
    if a.ndim == 0:
        return a.reshape(1, 1)
    if a.ndim == 1:
        return a.reshape(1, -1)
    return a


def _is_scalar(a):
    return a.size == 1


def _pad_dims(a, n):
    return a.reshape(a.shape + (1,) * (n - a.ndim)) if a.ndim < n else a


def _bcast(a, b):
    n = max(a.ndim, b.ndim)
    return _pad_dims(a, n), _pad_dims(b, n)


def _trim(a):
    """drop trailing singleton dimensions beyond the second (MATLAB never shows them)"""
    while a.ndim > 2 and a.shape[-1] == 1:
        a = a.reshape(a.shape[:-1])
    return a


# --------------------------------------------------------------------------------------------- tokenizer
_TOKEN = re.compile(r"""
    (?P<num>(\d+(\.(?![*/^'])\d*)?|\.\d+)([eE][+-]?\d+)?([ij](?![A-Za-z0-9_]))?) |
    (?P<id>[A-Za-z_]\w*) |
    (?P<op>\.\*|\./|\.\^|\.'|==|~=|<=|>=|&&|\|\||[-+*/\\^'<>=~:,;()\[\]{}.@&|]) |
    (?P<str>"[^"]*") |
    (?P<ws>[ \t]+) |
    (?P<nl>\n)
""", re.X)


class Tok:
    __slots__ = ("kind", "text", "space_before", "space_after")

    def __init__(self, kind, text, space_before):
        self.kind, self.text, self.space_before, self.space_after = kind, text, space_before, False

    def __repr__(self):
        return "%s(%r)" % (self.kind, self.text)


def tokenize(src):
    # strip comments (a % outside a string) and join continuation lines
    lines = []
    for raw in src.splitlines():
        out, in_str, i = [], False, 0
        while i < len(raw):
            c = raw[i]
            if c == '"':
                in_str = not in_str
            if c == "%" and not in_str:
                break
            out.append(c)
            i += 1
        lines.append("".join(out))
    text = "\n".join(lines)
    text = re.sub(r"\.\.\.[^\n]*\n", " ", text)
    toks, pos, space = [], 0, False
    while pos < len(text):
        if text[pos] == "'":
            prev = toks[-1] if toks else None
            operand = prev is not None and not space and (prev.kind in ("id", "num") or prev.text in (")", "]", "}", "'", ".'"))
            if not operand:                       # a quote that does not follow an operand opens a char literal
                end = pos + 1
                while True:
                    end = text.index("'", end)
                    if text[end:end + 2] == "''":
                        end += 2
                        continue
                    break
                toks.append(Tok("str", text[pos + 1:end].replace("''", "'"), space))
                space = False
                pos = end + 1
                continue
        if text[pos] == '"':                      # string-array literal "..." (treated like a char vector here)
            end = text.index('"', pos + 1)
            toks.append(Tok("str", text[pos + 1:end], space))
            space = False
            pos = end + 1
            continue
        m = _TOKEN.match(text, pos)
        if not m:
            raise MatlabError("cannot tokenize at %r" % text[pos:pos + 30])
        pos = m.end()
        kind = m.lastgroup
        if kind == "ws":
            space = True
            if toks:
                toks[-1].space_after = True
            continue
        t = m.group(kind)
        # a quote right after an operand is a transpose, otherwise it opens a char literal (not needed here)
        toks.append(Tok(kind, t, space))
        space = False
    toks.append(Tok("nl", "\n", False))
    toks.append(Tok("eof", "", False))
    return toks


# --------------------------------------------------------------------------------------------- parser
class Parser:
    def __init__(self, toks):
        self.t, self.i = toks, 0
        self.in_matrix = 0
        self.in_index = 0

    def peek(self, k=0):
        return self.t[self.i + k]

    def next(self):
        tok = self.t[self.i]
        self.i += 1
        return tok

    def accept(self, text):
        if self.peek().text == text and self.peek().kind in ("op", "id"):
            return self.next()
        return None

    def expect(self, text):
        tok = self.next()
        if tok.text != text:
            raise MatlabError("expected %r, got %r" % (text, tok.text))
        return tok

    def skip_newlines(self):
        while self.peek().kind == "nl" or self.peek().text in (";", ","):
            self.next()

    # ---- file / statements
    def parse_file(self):
        funcs = {}
        self.skip_newlines()
        while self.peek().kind != "eof":
            if self.peek().text != "function":
                raise MatlabError("only function files are supported, got %r" % self.peek().text)
            f = self.parse_function()
            funcs.setdefault(f["name"], f)
            self.skip_newlines()
        return funcs

    def parse_function(self):
        self.expect("function")
        outs = []
        # forms: function name(args) | function out = name(args) | function [o1,o2] = name(args)
        if self.peek().text == "[":
            self.next()
            while self.peek().text != "]":
                if self.peek().text == ",":
                    self.next()
                    continue
                outs.append(self.next().text)
            self.next()
            self.expect("=")
            name = self.next().text
        else:
            first = self.next().text
            if self.accept("="):
                outs, name = [first], self.next().text
            else:
                name = first
        args = []
        if self.accept("("):
            while self.peek().text != ")":
                if self.peek().text == ",":
                    self.next()
                    continue
                args.append(self.next().text)
            self.next()
        body = self.parse_block(("end", "function"))
        if self.peek().text == "end":
            self.next()
        return dict(name=name, args=args, outs=outs, body=body)

    def parse_block(self, terminators):
        stmts = []
        while True:
            self.skip_newlines()
            tok = self.peek()
            if tok.kind == "eof" or (tok.kind == "id" and tok.text in terminators):
                return stmts
            stmts.append(self.parse_statement())

    def parse_statement(self):
        tok = self.peek()
        if tok.kind == "id" and tok.text == "for":
            self.next()
            paren = self.accept("(")
            var = self.next().text
            self.expect("=")
            rng = self.parse_expr()
            if paren:
                self.expect(")")
            body = self.parse_block(("end",))
            self.expect("end")
            return ("for", var, rng, body)
        if tok.kind == "id" and tok.text == "while":
            self.next()
            cond = self.parse_expr()
            body = self.parse_block(("end",))
            self.expect("end")
            return ("while", cond, body)
        if tok.kind == "id" and tok.text == "if":
            self.next()
            branches, other = [], None
            cond = self.parse_expr()
            body = self.parse_block(("end", "else", "elseif"))
            branches.append((cond, body))
            while True:
                if self.accept("elseif"):
                    cond = self.parse_expr()
                    branches.append((cond, self.parse_block(("end", "else", "elseif"))))
                elif self.accept("else"):
                    other = self.parse_block(("end",))
                else:
                    break
            self.expect("end")
            return ("if", branches, other)
        # multi-output assignment  [a, ~, c] = f(...)
        if tok.text == "[":
            save = self.i
            try:
                self.next()
                names = []
                while self.peek().text != "]":
                    if self.peek().text == ",":
                        self.next()
                        continue
                    t = self.next()
                    if t.kind != "id" and t.text != "~":
                        raise MatlabError("not an lvalue list")
                    names.append(t.text)
                self.next()
                if self.peek().text == "=" and self.peek(1).text != "=":
                    self.next()
                    rhs = self.parse_expr()
                    return ("massign", names, rhs)
                raise MatlabError("not an assignment")
            except MatlabError:
                self.i = save
        # assignment or expression statement
        save = self.i
        if tok.kind == "id":
            lhs = self.parse_postfix()
            if self.peek().text == "=" and self.peek().kind == "op":
                self.next()
                rhs = self.parse_expr()
                return ("assign", lhs, rhs)
            self.i = save
        return ("expr", self.parse_expr())

    # ---- expressions (MATLAB precedence, lowest first)
    def parse_expr(self):
        return self.parse_oror()

    def parse_oror(self):
        a = self.parse_andand()
        while self.peek().text == "||":
            self.next()
            a = ("bin", "||", a, self.parse_andand())
        return a

    def parse_andand(self):
        a = self.parse_cmp()
        while self.peek().text == "&&":
            self.next()
            a = ("bin", "&&", a, self.parse_cmp())
        return a

    def parse_cmp(self):
        a = self.parse_range()
        while self.peek().text in ("==", "~=", "<", ">", "<=", ">=") and self.peek().kind == "op":
            op = self.next().text
            a = ("bin", op, a, self.parse_range())
        return a

    def parse_range(self):
        a = self.parse_add()
        if self.peek().text == ":" and self.peek().kind == "op" and not self._colon_is_bare():
            self.next()
            b = self.parse_add()
            if self.peek().text == ":" and self.peek().kind == "op":
                self.next()
                c = self.parse_add()
                return ("range", a, b, c)
            return ("range", a, None, b)
        return a

    def _colon_is_bare(self):
        return False

    def _binary_in_matrix(self, tok):
        """inside [ ], 'a -b' is two elements while 'a - b' and 'a-b' are one"""
        return not (self.in_matrix and not self.in_index_depth() and tok.space_before and not tok.space_after)

    def in_index_depth(self):
        return self.in_index > 0

    def parse_add(self):
        a = self.parse_mul()
        while self.peek().text in ("+", "-") and self.peek().kind == "op" and self._binary_in_matrix(self.peek()):
            op = self.next().text
            a = ("bin", op, a, self.parse_mul())
        return a

    def parse_mul(self):
        a = self.parse_unary()
        while self.peek().text in ("*", "/", "\\", ".*", "./") and self.peek().kind == "op":
            op = self.next().text
            a = ("bin", op, a, self.parse_unary())
        return a

    def parse_unary(self):
        tok = self.peek()
        if tok.kind == "op" and tok.text in ("-", "+", "~"):
            self.next()
            return ("un", tok.text, self.parse_unary())
        return self.parse_power()

    def parse_power(self):
        a = self.parse_postfix()
        while self.peek().text in ("^", ".^") and self.peek().kind == "op":
            op = self.next().text
            tok = self.peek()
            if tok.kind == "op" and tok.text in ("-", "+"):
                self.next()
                b = ("un", tok.text, self.parse_postfix())
            else:
                b = self.parse_postfix()
            a = ("bin", op, a, b)
        return a

    def parse_postfix(self):
        a = self.parse_primary()
        while True:
            tok = self.peek()
            if tok.text == "(" and not (self.in_matrix and not self.in_index_depth() and tok.space_before):
                self.next()
                self.in_index += 1
                args = self.parse_args(")")
                self.in_index -= 1
                a = ("call", a, args)
            elif tok.text == "." and self.peek(1).kind == "id" and not tok.space_before:
                self.next()
                a = ("field", a, self.next().text)
            elif tok.text in ("'", ".'") and tok.kind == "op" and not tok.space_before:
                self.next()
                a = ("un", tok.text, a)
            else:
                return a

    def parse_args(self, closer):
        args = []
        while self.peek().text != closer:
            if self.peek().text == ",":
                self.next()
                continue
            if self.peek().text == ":" and self.peek(1).text in (",", closer):
                self.next()
                args.append(("colon",))
            else:
                args.append(self.parse_expr())
        self.next()
        return args

    def parse_primary(self):
        tok = self.next()
        if tok.kind == "str":
            return ("str", tok.text.strip('"'))
        if tok.kind == "num":
            if tok.text[-1] in "ij":
                return ("num", 1j * float(tok.text[:-1]))
            return ("num", float(tok.text))
        if tok.kind == "id":
            if tok.text == "end" and self.in_index:
                return ("endidx",)
            return ("id", tok.text)
        if tok.text == "(":
            saved_m, saved_i = self.in_matrix, self.in_index
            self.in_matrix, self.in_index = 0, 0
            e = self.parse_expr()
            self.in_matrix, self.in_index = saved_m, saved_i
            self.expect(")")
            return ("paren", e)
        if tok.text == "[":
            saved_i = self.in_index
            self.in_matrix += 1
            self.in_index = 0
            rows, row = [], []
            while True:
                t = self.peek()
                if t.text == "]":
                    self.next()
                    break
                if t.text == ";" or t.kind == "nl":
                    self.next()
                    if row:
                        rows.append(row)
                    row = []
                    continue
                if t.text == ",":
                    self.next()
                    continue
                row.append(self.parse_expr())
            if row:
                rows.append(row)
            self.in_matrix -= 1
            self.in_index = saved_i
            return ("matrix", rows)
        raise MatlabError("unexpected token %r" % tok.text)


# --------------------------------------------------------------------------------------------- evaluator
class _Struct(dict):
    pass


def _is_colon(idx):
    return isinstance(idx, tuple) and idx == ("colon",)


def _index_list(idx, n):
    """MATLAB subscript -> 0-based numpy index array for a dimension of extent n"""
    if _is_colon(idx):
        return np.arange(n)
    a = np.asarray(idx)
    if a.dtype == bool:
        return np.flatnonzero(a.ravel(order="F"))
    r = np.rint(a.ravel(order="F").real).astype(np.int64)
    if np.any(np.abs(a.ravel(order="F").real - r) > 0) or np.any(r < 1):
        raise MatlabError("subscript indices must be positive integers")
    return r - 1


class MatlabFile:
    """functions of one or more .m sources; call with ``f.call(name, args, nargout)``"""

    def __init__(self, *sources, externals=None):
        self.funcs = {}
        for src in sources:
            for k, v in Parser(tokenize(src)).parse_file().items():
                self.funcs.setdefault(k, v)
        self.ext = dict(externals or {})

    # ---- builtins
    def _builtin(self, name, args, nargout):
        A = [a for a in args]
        if name == "size":
            x = A[0]
            if len(A) == 2:
                d = int(A[1].item())
                return [mat(float(x.shape[d - 1] if d <= x.ndim else 1))]
            shp = list(x.shape)
            if nargout <= 1:
                return [mat(np.array(shp, dtype=float))]
            if nargout < len(shp):
                shp = shp[:nargout - 1] + [int(np.prod(shp[nargout - 1:]))]
            shp += [1] * (nargout - len(shp))
            return [mat(float(s)) for s in shp]
        if name in ("zeros", "ones"):
            dims = [int(a.item()) for a in A if not isinstance(a, str)] or [1]
            if len(dims) == 1:
                dims = dims * 2
            return [np.zeros(dims) if name == "zeros" else np.ones(dims)]
        if name == "eye":
            dims = [int(a.item()) for a in A]
            return [np.eye(dims[0], dims[-1])]
        if name == "complex":
            return [A[0].astype(np.complex128) if len(A) == 1 else A[0] + 1j * A[1]]
        if name == "numel":
            return [mat(float(A[0].size))]
        if name == "length":
            return [mat(float(max(A[0].shape) if A[0].size else 0))]
        if name == "squeeze":
            x = A[0]
            if x.ndim <= 2:
                return [x]
            shp = [s for s in x.shape if s != 1]
            while len(shp) < 2:
                shp.append(1)
            return [x.reshape(shp, order="F")]
        if name == "reshape":
            dims = [int(v) for v in (A[1].ravel() if len(A) == 2 else [a.item() for a in A[1:]])]
            return [mat(A[0].reshape(dims, order="F"))]
        if name == "repmat":
            reps = [int(a.item()) for a in A[1:]] if len(A) > 2 else [int(v) for v in A[1].ravel()]
            if len(reps) == 1:
                reps = reps * 2
            return [np.tile(_pad_dims(A[0], len(reps)), reps)]
        if name == "inv":
            return [np.linalg.inv(A[0])]
        if name == "sum":
            x = A[0]
            if len(A) == 2:
                return [np.sum(x, axis=int(A[1].item()) - 1, keepdims=True)]
            ax = next((i for i, s in enumerate(x.shape) if s != 1), 0)
            return [np.sum(x, axis=ax, keepdims=True)]
        if name == "sqrt":
            x = A[0]
            return [np.sqrt(x.astype(np.complex128)) if np.iscomplexobj(x) or np.any(x.real < 0) else np.sqrt(x)]
        simple = {"conj": np.conj, "abs": np.abs, "real": np.real, "imag": np.imag, "floor": np.floor, "exp": np.exp,
                  "round": np.round, "double": lambda v: v.astype(np.complex128 if np.iscomplexobj(v) else np.float64)}
        if name in simple:
            return [mat(simple[name](A[0]))]
        if name == "transpose":
            return [np.swapaxes(A[0], 0, 1)]
        if name == "ctranspose":
            return [np.conj(np.swapaxes(A[0], 0, 1))]
        if name == "max" and len(A) == 1:
            x = A[0]
            v = x.ravel(order="F") if min(x.shape) == 1 else None
            if v is None:
                raise MatlabError("max: only vectors are supported")
            key = np.abs(v) if np.iscomplexobj(v) else v            # MATLAB compares complex numbers by magnitude
            k = int(np.argmax(key))                                 # first maximum
            return [mat(v[k]), mat(float(k + 1))]
        if name == "diag" and len(A) == 1:
            x = A[0]
            if min(x.shape) == 1:
                return [np.diag(x.ravel(order="F"))]
            return [np.diag(x).reshape(-1, 1)]
        if name == "norm" and len(A) == 2 and isinstance(A[1], str):
            if A[1] != "fro":
                raise MatlabError("norm: unsupported option %r" % A[1])
            return [mat(np.sqrt(np.sum(np.abs(A[0]) ** 2)))]
        if name == "norm":
            x = A[0]
            if min(x.shape) == 1 or x.ndim > 2:
                return [mat(np.sqrt(np.sum(np.abs(x) ** 2)))]
            return [mat(np.linalg.norm(x, 2))]
        if name == "mean":
            x = A[0]
            if len(A) == 2 and isinstance(A[1], str):
                if A[1] != "all":
                    raise MatlabError("mean: unsupported option %r" % A[1])
                return [mat(np.mean(x))]
            if len(A) == 2:
                return [np.mean(x, axis=int(A[1].item()) - 1, keepdims=True)]
            ax = next((i for i, s in enumerate(x.shape) if s != 1), 0)
            return [np.mean(x, axis=ax, keepdims=True)]
        if name == "mod":
            return [np.mod(A[0], A[1])]
        if name == "isempty":
            return [mat(float(A[0].size == 0))]
        if name in ("num2str", "int2str"):
            v = A[0] if isinstance(A[0], str) else np.asarray(A[0]).real.item()
            return [v if isinstance(v, str) else (str(int(round(v))) if float(v).is_integer() or name == "int2str" else "%.4g" % v)]
        if name == "strcat":
            return ["".join(a if isinstance(a, str) else str(a) for a in A)]
        if name == "insertBefore":              # insertBefore(str, pat, new): before every occurrence of pat
            return [A[0].replace(A[1], A[2] + A[1])]
        if name == "disp":
            return [np.zeros((0, 0))]
        if name == "pi":
            return [mat(np.pi)]
        if name in ("eps",):
            return [mat(np.finfo(float).eps)]
        return None

    # ---- evaluation
    def call(self, name, args, nargout=1):
        args = [a if isinstance(a, (_Struct, str)) else (a if isinstance(a, dict) else mat(a)) for a in args]
        args = [_Struct(a) if isinstance(a, dict) and not isinstance(a, _Struct) else a for a in args]
        if name in self.funcs:
            f = self.funcs[name]
            env = {}
            for k, v in zip(f["args"], args):
                env[k] = v
            env["__nargin__"] = len(args)
            self._exec_block(f["body"], env)
            outs = []
            for o in f["outs"][:max(1, nargout)]:
                if o not in env:
                    raise MatlabError("output %r of %s not assigned" % (o, name))
                outs.append(env[o])
            return outs
        if name in self.ext:
            r = self.ext[name](*args)
            r = list(r) if isinstance(r, (tuple, list)) else [r]
            return [mat(v) for v in r]
        b = self._builtin(name, args, nargout)
        if b is None:
            raise MatlabError("undefined function %r (supply it through externals=)" % name)
        return b

    def _exec_block(self, stmts, env):
        for s in stmts:
            kind = s[0]
            if kind == "assign":
                self._assign(s[1], self._eval(s[2], env), env)
            elif kind == "massign":
                vals = self._eval_multi(s[2], env, len(s[1]))
                for nm, v in zip(s[1], vals):
                    if nm != "~":
                        env[nm] = v
            elif kind == "expr":
                if s[1][0] == "call" and s[1][1] == ("id", "load") and "load" not in env:
                    # load(file): every variable of the MAT-file lands in the workspace (structs keep their fields)
                    self._load_into(env, self._eval(s[1][2][0], env))
                else:
                    self._eval(s[1], env)
            elif kind == "for":
                rng = self._eval(s[2], env)
                cols = rng.reshape(rng.shape[0], -1, order="F")
                for c in range(cols.shape[1]):
                    env[s[1]] = mat(cols[:, c].reshape(-1, 1) if cols.shape[0] > 1 else cols[0, c])
                    self._exec_block(s[3], env)
            elif kind == "while":
                guard = 0
                while self._truth(self._eval(s[1], env)):
                    self._exec_block(s[2], env)
                    guard += 1
                    if guard > 100000:
                        raise MatlabError("while loop does not terminate")
            elif kind == "if":
                done = False
                for cond, body in s[1]:
                    if self._truth(self._eval(cond, env)):
                        self._exec_block(body, env)
                        done = True
                        break
                if not done and s[2] is not None:
                    self._exec_block(s[2], env)

    @staticmethod
    def _load_into(env, path):
        from scipy.io import loadmat
        for k, v in loadmat(path, struct_as_record=False, squeeze_me=False).items():
            if k.startswith("__"):
                continue
            if isinstance(v, np.ndarray) and v.dtype == object and v.size == 1 and hasattr(v.flat[0], "_fieldnames"):
                st = v.flat[0]
                env[k] = _Struct({f: mat(np.asarray(getattr(st, f), dtype=np.complex128 if np.iscomplexobj(getattr(st, f)) else np.float64))
                                  for f in st._fieldnames})
            else:
                env[k] = mat(np.asarray(v, dtype=np.complex128 if np.iscomplexobj(v) else np.float64))

    @staticmethod
    def _truth(v):
        v = np.asarray(v)
        return v.size > 0 and bool(np.all(v != 0))

    def _eval_multi(self, node, env, nargout):
        if node[0] == "call" and node[1][0] == "id" and node[1][1] not in env:
            args = [self._eval_arg(a, env, None, 0, 1) for a in node[2]]
            outs = self.call(node[1][1], args, nargout)
            if len(outs) < nargout:
                raise MatlabError("too many output arguments for %s" % node[1][1])
            return outs
        return [self._eval(node, env)]

    def _eval_arg(self, a, env, base, dim, ndims):
        if _is_colon(a):
            return a
        if base is not None:
            # value of `end` in this subscript position
            if ndims == 1:
                end = base.size
            else:
                shp = list(base.shape) + [1] * max(0, ndims - base.ndim)
                end = shp[dim] if dim < ndims - 1 else int(np.prod(shp[dim:]))
            env = dict(env)
            env["__end__"] = mat(float(end))
        return self._eval(a, env)

    def _eval(self, node, env):
        k = node[0]
        if k == "num":
            return mat(node[1])
        if k == "str":
            return node[1]
        if k == "paren":
            return self._eval(node[1], env)
        if k == "endidx":
            return env["__end__"]
        if k == "id":
            name = node[1]
            if name in env:
                return env[name]
            if name in ("i", "j", "1i", "1j"):
                return mat(1j)
            if name == "pi":
                return mat(np.pi)
            if name == "nargin":
                return mat(float(env["__nargin__"]))
            if name in ("true", "false"):
                return mat(1.0 if name == "true" else 0.0)
            return self.call(name, [], 1)[0]
        if k == "field":
            base = self._eval(node[1], env)
            if not isinstance(base, dict) or node[2] not in base:
                raise MatlabError("no field %r" % node[2])
            v = base[node[2]]
            return v if isinstance(v, dict) else mat(v)
        if k == "matrix":
            rows = []
            for r in node[1]:
                vals = [self._eval(e, env) for e in r]
                vals = [v for v in vals if v.size > 0 or len(vals) == 1]
                rows.append(np.concatenate(vals, axis=1) if len(vals) > 1 else vals[0])
            if not rows:
                return np.zeros((0, 0))
            return np.concatenate(rows, axis=0) if len(rows) > 1 else rows[0]
        if k == "range":
            a = self._eval(node[1], env).real.item()
            b = self._eval(node[3], env).real.item()
            step = 1.0 if node[2] is None else self._eval(node[2], env).real.item()
            n = int(np.floor((b - a) / step + 1e-10)) + 1
            return mat(a + step * np.arange(max(n, 0))) if n > 0 else np.zeros((1, 0))
        if k == "un":
            v = self._eval(node[2], env)
            if node[1] == "-":
                return -v
            if node[1] == "+":
                return v
            if node[1] == "~":
                return mat((v == 0).astype(float))
            if node[1] == "'":
                return np.conj(np.swapaxes(v, 0, 1))
            if node[1] == ".'":
                return np.swapaxes(v, 0, 1)
        if k == "bin":
            op = node[1]
            if op == "&&":
                return mat(float(self._truth(self._eval(node[2], env)) and self._truth(self._eval(node[3], env))))
            if op == "||":
                return mat(float(self._truth(self._eval(node[2], env)) or self._truth(self._eval(node[3], env))))
            a, b = self._eval(node[2], env), self._eval(node[3], env)
            if op in ("+", "-", ".*", "./", ".^", "==", "~=", "<", ">", "<=", ">="):
                a, b = _bcast(a, b)
                f = {"+": np.add, "-": np.subtract, ".*": np.multiply, "./": np.divide, ".^": np.power,
                     "==": np.equal, "~=": np.not_equal, "<": np.less, ">": np.greater, "<=": np.less_equal,
                     ">=": np.greater_equal}[op]
                r = f(a, b)
                return r.astype(float) if r.dtype == bool else r
            if op == "*":
                return a * b if (_is_scalar(a) or _is_scalar(b)) else a @ b
            if op == "/":
                if _is_scalar(b):
                    return a / b
                return np.linalg.solve(b.T, a.T).T                # A / B = A * inv(B)
            if op == "\\":                                      # A \ B: square A -> LU solve (MATLAB's mldivide)
                if _is_scalar(a):
                    return b / a
                if a.shape[0] != a.shape[1]:
                    raise MatlabError("mldivide: only square systems are supported")
                return np.linalg.solve(a, b)
            if op == "^":
                if _is_scalar(a) and _is_scalar(b):
                    return np.power(a.astype(np.complex128) if (a.real < 0).any() else a, b)
                return np.linalg.matrix_power(a, int(b.item()))
        if k == "call":
            target = node[1]
            if target[0] == "id" and target[1] not in env:
                args = [self._eval_arg(a, env, None, 0, 1) for a in node[2]]
                return self.call(target[1], args, 1)[0]
            base = self._eval(target, env)
            return self._index(base, node[2], env)
        raise MatlabError("cannot evaluate %r" % (node,))

    def _index(self, base, argnodes, env):
        n = len(argnodes)
        if n == 0:
            return base
        idx = [self._eval_arg(a, env, base, d, n) for d, a in enumerate(argnodes)]
        if n == 1:
            flat = base.ravel(order="F")
            if _is_colon(idx[0]):
                return flat.reshape(-1, 1)
            ii = _index_list(idx[0], flat.size)
            shape_src = np.asarray(idx[0])
            out = flat[ii]
            if shape_src.ndim >= 2 and min(shape_src.shape) > 1:
                return out.reshape(shape_src.shape, order="F")
            # vector source and vector index: orientation of the source; otherwise of the index
            if min(base.shape) == 1 and base.ndim == 2 and base.size > 1:
                return out.reshape(-1, 1) if base.shape[1] == 1 else out.reshape(1, -1)
            return out.reshape(shape_src.shape if shape_src.ndim >= 2 else (1, -1))
        shp = list(base.shape) + [1] * max(0, n - base.ndim)
        if n < len(shp):
            shp = shp[:n - 1] + [int(np.prod(shp[n - 1:]))]
        b = base.reshape(shp, order="F")
        lists = [_index_list(ix, shp[d]) for d, ix in enumerate(idx)]
        return mat(_trim(b[np.ix_(*lists)]))

    def _assign(self, lhs, val, env):
        if lhs[0] == "id":
            env[lhs[1]] = val
            return
        if lhs[0] == "field":
            base = lhs[1]
            if base[0] != "id":
                raise MatlabError("nested struct assignment is not supported")
            env.setdefault(base[1], _Struct())[lhs[2]] = val
            return
        if lhs[0] == "call" and lhs[1][0] == "id":
            name = lhs[1][1]
            cur = env.get(name, np.zeros((0, 0)))
            n = len(lhs[2])
            idx = [self._eval_arg(a, env, cur, d, n) for d, a in enumerate(lhs[2])]
            if np.iscomplexobj(val) and not np.iscomplexobj(cur):
                cur = cur.astype(np.complex128)
            if n == 1:
                flat = cur.ravel(order="F").copy()
                ii = _index_list(idx[0], flat.size)
                if ii.size and ii.max() >= flat.size:
                    grown = np.zeros(ii.max() + 1, dtype=flat.dtype)
                    grown[:flat.size] = flat
                    flat = grown
                    cur = flat.reshape(1, -1) if cur.shape[0] <= 1 else flat.reshape(-1, 1)
                flat[ii] = val.ravel(order="F") if val.size > 1 else val.item()
                env[name] = flat.reshape(cur.shape, order="F")
                return
            shp = list(cur.shape) + [1] * max(0, n - cur.ndim)
            if cur.size == 0:
                # A(:,j,i) = v on an undefined / empty A: the colon dimension takes its extent from v
                shp = [0] * n
                fixed = int(np.prod([np.asarray(ix).size for ix in idx if not _is_colon(ix)])) or 1
                colons = [d for d, ix in enumerate(idx) if _is_colon(ix)]
                for d in colons:
                    shp[d] = val.size // fixed if len(colons) == 1 else val.shape[d] if d < val.ndim else 1
                cur = np.zeros(shp, dtype=val.dtype)
            lists = []
            for d, ix in enumerate(idx):
                lists.append(_index_list(ix, shp[d] if d < len(shp) else 1))
            need = [max(shp[d], (int(l.max()) + 1) if l.size else 0) for d, l in enumerate(lists)]
            if need != shp[:n]:
                grown = np.zeros(need + shp[n:], dtype=cur.dtype)
                grown[tuple(slice(0, s) for s in shp)] = cur.reshape(shp)
                cur, shp = grown, list(grown.shape)
            else:
                cur = cur.reshape(shp).copy()
            target_shape = [len(l) for l in lists]
            v = val if val.size == 1 else val.reshape([s for s in target_shape], order="F") \
                if val.size == int(np.prod(target_shape)) else val
            cur[np.ix_(*lists)] = v.item() if val.size == 1 else v
            env[name] = mat(_trim(cur))
            return
        raise MatlabError("unsupported assignment target %r" % (lhs,))
