#!/usr/bin/env python3
"""vl2c - a small Verilog-2001 -> C++ translator (cycle based, 2-state) for the subset of the language
that /root/reference/RTL/mpeg2encoder.v uses.  TEST INFRASTRUCTURE.

Purpose: neither the build image nor the GPU box has a Verilog simulator, so the reference cannot be
run as its authors run it (SIM/tb_run_iverilog.bat).  This tool reads the reference RTL WHERE IT LIES
(nothing of it is copied into the repository), elaborates it for one parameter set and writes a C++
model into oracle/_ref/ (git-ignored) which oracle/rtl_tb.cpp - a C++ restatement of the stimulus of
SIM/tb_mpeg2encoder.v - drives.  The result is the reference's own logic executed clock by clock; it is
used to pin oracle/m2v_oracle.c (tests/test_rtl_pin.py) and as the 'reference' CPU baseline.

Semantics implemented (IEEE 1364-2001): expression sizing and signedness (self-determined /
context-determined operands, sign of an expression = all operands signed, extension by the propagated
type), blocking vs non-blocking assignment (shadow copies committed after every clocked block has run),
continuous assignments and always@(*) evaluated in source order before and after each edge, functions,
multi-dimensional arrays with arbitrary (negative) bounds, part selects (+:), concatenations on both
sides, case, for, integer variables.  2-state: X/Z do not exist, uninitialised storage is 0 (the RTL's
outputs do not depend on uninitialised storage: SURVEY.md 3.4).  Values wider than 64 bits use a
fixed 512-bit helper type.  Not supported (unused by the RTL): generate, tasks, delays, events other
than posedge clk / negedge rstn, inout, tri-state, real, strings.

usage: vl2c.py <rtl.v> <out.cpp> PARAM=value ...
"""
import re
import sys

# ------------------------------------------------------------------------------------------------
# lexer
# ------------------------------------------------------------------------------------------------
TOK = re.compile(r"""
    (?P<ws>\s+|//[^\n]*|/\*.*?\*/)
  | (?P<num>(?:\d[\d_]*)?\s*'[sS]?[bBdDhHoO]\s*[0-9a-fA-FxXzZ_]+|\d[\d_]*)
  | (?P<id>[A-Za-z_][A-Za-z0-9_$]*|\$[a-z]+)
  | (?P<op><<<|>>>|<<|>>|<=|>=|==|!=|&&|\|\||\+:|-:|[-+*/%<>!~&|^?:;,.=(){}\[\]@\#])
""", re.X | re.S)


def lex(text):
    out, i = [], 0
    while i < len(text):
        m = TOK.match(text, i)
        if not m:
            raise SyntaxError('lex error at %r' % text[i:i + 40])
        i = m.end()
        if m.lastgroup != 'ws':
            out.append((m.lastgroup, m.group(m.lastgroup)))
    out.append(('eof', ''))
    return out


def parse_number(s):
    """-> (value, width or None, signed)"""
    s = s.replace('_', '').replace(' ', '')
    if "'" not in s:
        return int(s), None, True                      # unsized decimal: 32-bit signed
    w, rest = s.split("'")
    signed = rest[0] in 'sS'
    if signed:
        rest = rest[1:]
    base = {'b': 2, 'd': 10, 'h': 16, 'o': 8}[rest[0].lower()]
    v = int(rest[1:].lower().replace('x', '0').replace('z', '0'), base)
    width = int(w) if w else None
    if width is not None:
        v &= (1 << width) - 1
    return v, width, signed


# ------------------------------------------------------------------------------------------------
# parser (recursive descent)
# ------------------------------------------------------------------------------------------------
class Parser:
    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self, k=0): return self.t[self.i + k]
    def at(self, v): return self.t[self.i][1] == v and self.t[self.i][0] != 'eof'
    def next(self): self.i += 1; return self.t[self.i - 1]

    def eat(self, v):
        if not self.at(v):
            raise SyntaxError('expected %r got %r near token %d (%s)' % (v, self.peek(), self.i, ' '.join(x[1] for x in self.t[max(0, self.i - 8):self.i + 4])))
        return self.next()

    def opt(self, v):
        if self.at(v):
            self.next(); return True
        return False

    # ---- expressions (precedence climbing) ----
    BIN = [['||'], ['&&'], ['|'], ['^'], ['&'], ['==', '!='], ['<', '<=', '>', '>='], ['<<', '>>', '<<<', '>>>'], ['+', '-'], ['*', '/', '%']]

    def expr(self): return self.cond()

    def cond(self):
        c = self.binary(0)
        if self.opt('?'):
            a = self.cond(); self.eat(':'); b = self.cond()
            return ('cond', c, a, b)
        return c

    def binary(self, lvl):
        if lvl == len(self.BIN):
            return self.unary()
        a = self.binary(lvl + 1)
        while self.peek()[0] == 'op' and self.peek()[1] in self.BIN[lvl]:
            op = self.next()[1]
            b = self.binary(lvl + 1)
            a = ('bin', op, a, b)
        return a

    def unary(self):
        if self.peek()[0] == 'op' and self.peek()[1] in ('-', '+', '!', '~', '&', '|', '^'):
            op = self.next()[1]
            return ('un', op, self.unary())
        return self.postfix(self.primary())

    def primary(self):
        k, v = self.peek()
        if k == 'num':
            self.next(); return ('num',) + parse_number(v)
        if v == '(':
            self.next(); e = self.expr(); self.eat(')'); return e
        if v == '{':
            self.next()
            first = self.expr()
            if self.at('{'):                            # replication {n{...}}
                self.next(); parts = [self.expr()]
                while self.opt(','): parts.append(self.expr())
                self.eat('}'); self.eat('}')
                return ('repl', first, ('concat', parts))
            parts = [first]
            while self.opt(','): parts.append(self.expr())
            self.eat('}')
            return ('concat', parts)
        if k == 'id':
            self.next()
            if v in ('$signed', '$unsigned'):
                self.eat('('); e = self.expr(); self.eat(')')
                return ('sys', v, e)
            if self.at('(') and not v.startswith('$'):
                self.next(); args = []
                if not self.at(')'):
                    args.append(self.expr())
                    while self.opt(','): args.append(self.expr())
                self.eat(')')
                return ('call', v, args)
            return ('id', v)
        raise SyntaxError('unexpected %r' % (self.peek(),))

    def postfix(self, e):
        while self.at('['):
            self.next()
            a = self.expr()
            if self.opt(':'):
                b = self.expr(); self.eat(']'); e = ('range', e, a, b)
            elif self.opt('+:'):
                b = self.expr(); self.eat(']'); e = ('prange', e, a, b, +1)
            elif self.opt('-:'):
                b = self.expr(); self.eat(']'); e = ('prange', e, a, b, -1)
            else:
                self.eat(']'); e = ('index', e, a)
        return e

    # ---- declarations ----
    def rng(self):
        self.eat('['); a = self.expr(); self.eat(':'); b = self.expr(); self.eat(']')
        return (a, b)

    def decl_names(self, kind, signed, rg, out):
        while True:
            name = self.next()[1]
            dims = []
            while self.at('['):
                dims.append(self.rng())
            init = None
            if self.opt('='):
                init = self.expr()
            out.append(('decl', kind, name, signed, rg, dims, init))
            if not self.opt(','):
                break

    def module(self):
        items = []
        self.eat('module'); self.name = self.next()[1]
        if self.opt('#'):
            self.eat('(')
            while not self.at(')'):
                self.eat('parameter'); n = self.next()[1]; self.eat('='); items.append(('param', n, False, None, self.expr(), True)); self.opt(',')
            self.eat(')')
        self.eat('(')
        while not self.at(')'):
            d = self.next()[1]                           # input / output
            if self.at('wire') or self.at('reg'): self.next()
            signed = self.opt('signed')
            rg = self.rng() if self.at('[') else None
            while True:
                n = self.next()[1]
                items.append(('decl', d, n, signed, rg, [], None))
                if self.at(',') and self.peek(1)[1] not in ('input', 'output'):
                    self.next(); continue
                break
            self.opt(',')
        self.eat(')'); self.eat(';')
        while not self.at('endmodule'):
            items += self.item()
        return items

    def item(self):
        v = self.peek()[1]
        if v in ('localparam', 'parameter'):
            self.next(); signed = self.opt('signed'); rg = self.rng() if self.at('[') else None
            out = []
            while True:
                n = self.next()[1]; self.eat('='); out.append(('param', n, signed, rg, self.expr(), False))
                if not self.opt(','): break
            self.eat(';'); return out
        if v in ('wire', 'reg', 'integer'):
            self.next(); signed = self.opt('signed') or v == 'integer'
            rg = self.rng() if self.at('[') else None
            out = []; self.decl_names(v, signed, rg, out); self.eat(';'); return out
        if v == 'assign':
            self.next(); lhs = self.postfix(self.primary()); self.eat('='); rhs = self.expr(); self.eat(';')
            return [('assign', lhs, rhs)]
        if v == 'function':
            self.next(); signed = self.opt('signed'); rg = self.rng() if self.at('[') else None
            name = self.next()[1]; self.eat(';')
            decls = []
            while self.peek()[1] in ('input', 'reg', 'integer'):
                kind = self.next()[1]; s2 = self.opt('signed') or kind == 'integer'; r2 = self.rng() if self.at('[') else None
                self.decl_names(kind, s2, r2, decls); self.eat(';')
            body = self.stmt(); self.eat('endfunction')
            return [('function', name, signed, rg, decls, body)]
        if v == 'always':
            self.next(); self.eat('@'); self.eat('(')
            sens = []
            while not self.at(')'):
                sens.append(self.next()[1])
            self.eat(')')
            return [('always', sens, self.stmt())]
        raise SyntaxError('unsupported module item %r' % (self.peek(),))

    # ---- statements ----
    def stmt(self):
        v = self.peek()[1]
        if v == 'begin':
            self.next(); body = []
            while not self.at('end'): body.append(self.stmt())
            self.next(); return ('block', body)
        if v == 'if':
            self.next(); self.eat('('); c = self.expr(); self.eat(')')
            t = self.stmt(); e = None
            if self.opt('else'): e = self.stmt()
            return ('if', c, t, e)
        if v == 'case':
            self.next(); self.eat('('); sel = self.expr(); self.eat(')')
            arms, default = [], None
            while not self.at('endcase'):
                if self.opt('default'):
                    self.opt(':'); default = self.stmt(); continue
                labels = [self.expr()]
                while self.opt(','): labels.append(self.expr())
                self.eat(':'); arms.append((labels, self.stmt()))
            self.next(); return ('case', sel, arms, default)
        if v == 'for':
            self.next(); self.eat('(')
            init = self.simple_assign(); self.eat(';'); c = self.expr(); self.eat(';'); step = self.simple_assign(); self.eat(')')
            return ('for', init, c, step, self.stmt())
        if v == ';':
            self.next(); return ('block', [])
        s = self.simple_assign(); self.eat(';'); return s

    def simple_assign(self):
        lhs = self.postfix(self.primary())
        if self.opt('<='): return ('nba', lhs, self.expr())
        self.eat('='); return ('ba', lhs, self.expr())


# ------------------------------------------------------------------------------------------------
# elaboration + code generation
# ------------------------------------------------------------------------------------------------
class Sym:
    def __init__(self, kind, name, width, signed, lsb, dims, init=None):
        self.kind, self.name, self.width, self.signed, self.lsb, self.dims, self.init = kind, name, width, signed, lsb, dims, init
        self.size = 1
        for lo, hi in dims: self.size *= (hi - lo + 1)
        self.nba = False            # target of a non-blocking assignment somewhere
        self.cname = 'v_' + name


BIG = 64                            # values wider than this use the Big type


class Gen:
    def __init__(self, items, overrides):
        self.items, self.params, self.syms, self.funcs = items, {}, {}, {}
        self.overrides = overrides
        self.cur_func = None
        self.tmp = 0

    # ---- constant evaluation (parameters, ranges) ----
    def const(self, e):
        k = e[0]
        if k == 'num': return e[1] if not (e[3] and e[2] and (e[1] >> (e[2] - 1)) & 1) else e[1] - (1 << e[2])
        if k == 'id':
            if e[1] in self.params: return self.params[e[1]][0]
            raise ValueError('not a constant: ' + e[1])
        if k == 'un':
            v = self.const(e[2]); return {'-': -v, '+': v, '!': int(not v), '~': ~v}[e[1]]
        if k == 'bin':
            a, b = self.const(e[2]), self.const(e[3]); op = e[1]
            return {'+': a + b, '-': a - b, '*': a * b, '/': int(a / b) if b else 0, '%': a - b * int(a / b) if b else 0,
                    '<<': a << b if op == '<<' else 0, '>>': a >> b if op == '>>' else 0, '<': int(a < b), '>': int(a > b), '<=': int(a <= b), '>=': int(a >= b),
                    '==': int(a == b), '!=': int(a != b), '&&': int(bool(a) and bool(b)), '||': int(bool(a) or bool(b)),
                    '&': a & b, '|': a | b, '^': a ^ b}[op]
        if k == 'cond': return self.const(e[2]) if self.const(e[1]) else self.const(e[3])
        raise ValueError('constant expression kind ' + k)

    def elaborate(self):
        for it in self.items:
            if it[0] == 'param':
                _, n, signed, rg, val, is_port = it
                v = self.overrides[n] if (is_port and n in self.overrides) else self.const(val)
                if rg:
                    w = self.const(rg[0]) - self.const(rg[1]) + 1; s = signed
                else:
                    w, s = (val[2], val[3]) if (val[0] == 'num' and val[2]) else (32, True)
                self.params[n] = (v, w, s)
            elif it[0] == 'decl':
                _, kind, n, signed, rg, dims, init = it
                if rg: msb, lsb = self.const(rg[0]), self.const(rg[1])
                else: msb, lsb = (31, 0) if kind == 'integer' else (0, 0)
                ds = []
                for a, b in dims:
                    a, b = self.const(a), self.const(b); ds.append((min(a, b), max(a, b)))
                self.syms[n] = Sym(kind, n, msb - lsb + 1, signed, lsb, ds, init)
            elif it[0] == 'function':
                self.funcs[it[1]] = it
        # mark NBA targets
        def walk(s):
            if s[0] == 'block':
                for x in s[1]: walk(x)
            elif s[0] == 'if':
                walk(s[2]); s[3] and walk(s[3])
            elif s[0] == 'case':
                for _, b in s[2]: walk(b)
                s[3] and walk(s[3])
            elif s[0] == 'for':
                walk(s[4])
            elif s[0] == 'nba':
                for n in self.lhs_names(s[1]): self.syms[n].nba = True
        for it in self.items:
            if it[0] == 'always': walk(it[2])

    def lhs_names(self, l):
        if l[0] == 'id': return [l[1]]
        if l[0] in ('index', 'range', 'prange'): return self.lhs_names(l[1])
        if l[0] == 'concat': return sum((self.lhs_names(x) for x in l[1]), [])
        raise ValueError(l[0])

    # ---- symbol lookup (function locals shadow module symbols) ----
    def sym(self, n):
        if self.cur_func and n in self.cur_func: return self.cur_func[n]
        return self.syms.get(n)

    # ---- typing: self-determined (width, signed) ----
    def ty(self, e):
        k = e[0]
        if k == 'num': return (e[2] if e[2] else 32, e[3])
        if k == 'id':
            if self.sym(e[1]) is None and e[1] in self.params: return self.params[e[1]][1:]
            s = self.sym(e[1])
            if s is None: raise KeyError(e[1])
            return (s.width, s.signed)
        if k == 'index':
            base, depth = self.base_depth(e)
            s = self.sym(base)
            return (s.width, s.signed) if depth <= len(s.dims) else (1, False)
        if k == 'range': return (self.const(e[2]) - self.const(e[3]) + 1, False)
        if k == 'prange': return (self.const(e[3]), False)
        if k == 'un':
            if e[1] in ('!', '&', '|', '^'): return (1, False)
            return self.ty(e[2])
        if k == 'bin':
            op = e[1]
            if op in ('==', '!=', '<', '>', '<=', '>=', '&&', '||'): return (1, False)
            a = self.ty(e[2])
            if op in ('<<', '>>', '<<<', '>>>'): return a
            b = self.ty(e[3])
            return (max(a[0], b[0]), a[1] and b[1])
        if k == 'cond':
            a, b = self.ty(e[2]), self.ty(e[3]); return (max(a[0], b[0]), a[1] and b[1])
        if k == 'concat': return (sum(self.ty(x)[0] for x in e[1]), False)
        if k == 'repl': return (self.const(e[1]) * self.ty(e[2])[0], False)
        if k == 'call':
            f = self.funcs[e[1]]
            w = self.const(f[3][0]) - self.const(f[3][1]) + 1 if f[3] else 1
            return (w, f[2])
        if k == 'sys': return (self.ty(e[2])[0], e[1] == '$signed')
        raise ValueError(k)

    def base_depth(self, e):
        d = 0
        while e[0] == 'index': e = e[1]; d += 1
        return e[1], d

    # ---- C++ helpers ----
    @staticmethod
    def mask(w): return '0x%xull' % ((1 << w) - 1) if w < 64 else '~0ull'

    def ext(self, code, w, s, W, S):
        """pattern `code` of width w, sign s -> pattern of context width W under expression sign S"""
        if W > BIG or w > BIG:
            raise NotImplementedError
        if w >= W: return code if w == W else '((%s) & %s)' % (code, self.mask(W))
        if S and s: return '(SX(%s,%d) & %s)' % (code, w, self.mask(W))
        return code

    # ---- addressing ----
    def elem_index(self, e):
        """e = nested index over an array symbol -> (sym, [index exprs], leftover selects)"""
        chain = []
        while e[0] == 'index': chain.append(e[2]); e = e[1]
        chain.reverse()
        s = self.sym(e[1])
        return s, chain[:len(s.dims)], chain[len(s.dims):]

    def idx_code(self, s, idxs):
        """flattened element offset as C++ (int64)"""
        code = None
        for (lo, hi), ie in zip(s.dims, idxs):
            w, sg = self.ty(ie)
            v = self.rv(ie, w, sg)
            v = '((i64)SX(%s,%d))' % (v, w) if sg else '((i64)(%s))' % v
            term = '(%s - (%d))' % (v, lo)
            code = term if code is None else '(%s * %d + %s)' % (code, hi - lo + 1, term)
        return code

    def storage(self, s, shadow=False):
        return ('n_' if shadow else '') + s.cname

    # ---- rvalues ----
    def rv(self, e, W, S):
        """C++ expression (u64) = value of e evaluated in a context of width W and expression sign S, truncated to W"""
        if W > BIG: return self.rv_big(e, W, S) + '.lo64()'   # never happens for well-formed callers
        k = e[0]
        if k == 'num':
            w, s = self.ty(e)
            return self.ext('%dull' % e[1], w, s, W, S)
        if k == 'id':
            if self.sym(e[1]) is None and e[1] in self.params:
                v, w, s = self.params[e[1]]
                return self.ext('%dull' % (v & ((1 << w) - 1)), w, s, W, S)
            s = self.sym(e[1])
            if s.width > BIG: return '(%s.lo64() & %s)' % (s.cname, self.mask(W))
            return self.ext(s.cname, s.width, s.signed, W, S)
        if k == 'index':
            s, idxs, rest = self.elem_index(e)
            base = '%s' % s.cname if not s.dims else 'RD(%s,%s,%d)' % (s.cname, self.idx_code(s, idxs), s.size)
            if rest:                                         # bit select on the element
                assert len(rest) == 1
                bw, bs = self.ty(rest[0]); b = self.rv(rest[0], bw, bs)
                if s.width > BIG: return 'BITOF(%s,%s - %d)' % (base, b, s.lsb)
                return self.ext('((%s >> ((%s) - %d)) & 1ull)' % (base, b, s.lsb), 1, False, W, S)
            return self.ext(base, s.width, s.signed, W, S)
        if k in ('range', 'prange'):
            w, _ = self.ty(e)
            inner = e[1]
            if inner[0] == 'index': s, idxs, rest = self.elem_index(inner); assert not rest
            else: s, idxs = self.sym(inner[1]), []
            base = s.cname if not s.dims else 'RD(%s,%s,%d)' % (s.cname, self.idx_code(s, idxs), s.size)
            if k == 'range': lo = '%d' % (self.const(e[3]) - s.lsb)
            else:
                aw, asg = self.ty(e[2]); a = self.rv(e[2], aw, asg)
                lo = '((%s) - %d)' % (a, s.lsb) if e[4] > 0 else '((%s) - %d - %d)' % (a, s.lsb, w - 1)
            if s.width > BIG:
                assert w <= BIG; return self.ext('%s.slice(%s,%d)' % (base, lo, w), w, False, W, S)
            return self.ext('((%s >> (%s)) & %s)' % (base, lo, self.mask(w)), w, False, W, S)
        if k == 'un':
            op = e[1]
            if op == '!':
                w, s = self.ty(e[2]); return '(u64)(!(%s))' % self.nz(e[2])
            if op in ('&', '|', '^'):
                w, s = self.ty(e[2]); a = self.rv(e[2], w, s)
                return {'|': '(u64)((%s) != 0)' % a, '&': '(u64)((%s) == %s)' % (a, self.mask(w)), '^': '(u64)(__builtin_parityll(%s))' % a}[op]
            a = self.rv(e[2], W, S)
            if op == '+': return a
            return '((%s%s) & %s)' % ({'-': '0ull - ', '~': '~'}[op], a, self.mask(W))
        if k == 'bin':
            op = e[1]
            if op in ('&&', '||'): return '(u64)(%s %s %s)' % (self.nz(e[2]), op, self.nz(e[3]))
            if op in ('==', '!=', '<', '>', '<=', '>='):
                (wa, sa), (wb, sb) = self.ty(e[2]), self.ty(e[3]); w, s = max(wa, wb), sa and sb
                if w > BIG:
                    assert op in ('==', '!='); return '(u64)(%s(%s == %s))' % ('!' if op == '!=' else '', self.rv_big(e[2], w, s), self.rv_big(e[3], w, s))
                a, b = self.rv(e[2], w, s), self.rv(e[3], w, s)
                if s: return '(u64)((i64)SX(%s,%d) %s (i64)SX(%s,%d))' % (a, w, op, b, w)
                return '(u64)((%s) %s (%s))' % (a, op, b)
            if op in ('<<', '>>', '<<<', '>>>'):
                a = self.rv(e[2], W, S); bw, bs = self.ty(e[3]); b = self.rv(e[3], bw, False)
                if op in ('<<', '<<<'): return '(SHL(%s,%s) & %s)' % (a, b, self.mask(W))
                if op == '>>' or not S: return 'SHR(%s,%s)' % (a, b)
                return '(ASR(%s,%d,%s) & %s)' % (a, W, b, self.mask(W))
            a, b = self.rv(e[2], W, S), self.rv(e[3], W, S)
            if op in ('/', '%'):
                if S: return '((u64)SDIV%s((i64)SX(%s,%d),(i64)SX(%s,%d)) & %s)' % ('' if op == '/' else 'R', a, W, b, W, self.mask(W))
                return 'UDIV%s(%s,%s)' % ('' if op == '/' else 'R', a, b)
            return '((%s %s %s) & %s)' % (a, op, b, self.mask(W))
        if k == 'cond':
            return '(%s ? %s : %s)' % (self.nz(e[1]), self.rv(e[2], W, S), self.rv(e[3], W, S))
        if k in ('concat', 'repl'):
            w, _ = self.ty(e)
            if w > BIG: return '(%s.lo64() & %s)' % (self.rv_big(e, w, False), self.mask(W))
            parts = e[1] if k == 'concat' else [e[2]] * self.const(e[1])
            code, pos = [], w
            for p in parts:
                pw, ps = self.ty(p); pos -= pw
                code.append('(%s << %d)' % (self.rv(p, pw, ps), pos) if pos else self.rv(p, pw, ps))
            return self.ext('(' + ' | '.join(code) + ')', w, False, W, S)
        if k == 'call':
            w, s = self.ty(e)
            f = self.funcs[e[1]]
            ins = [d for d in f[4] if d[1] == 'input']
            args = []
            for d, a in zip(ins, e[2]):
                iw = self.const(d[4][0]) - self.const(d[4][1]) + 1 if d[4] else 1
                aw, asg = self.ty(a); cw = max(iw, aw)
                if cw > BIG: args.append(self.rv_big(a, cw, asg))
                else: args.append('((%s) & %s)' % (self.rv(a, cw, asg), self.mask(iw)))
            call = 'f_%s(%s)' % (e[1], ', '.join(args))
            if w > BIG: return '(%s.lo64() & %s)' % (call, self.mask(W))
            return self.ext(call, w, s, W, S)
        if k == 'sys':
            w, s = self.ty(e[2])
            if w > BIG: return '(%s.lo64() & %s)' % (self.rv_big(e[2], w, s), self.mask(W))
            return self.ext(self.rv(e[2], w, s), w, e[1] == '$signed', W, S)
        raise ValueError(k)

    def nz(self, e):
        w, s = self.ty(e)
        if w > BIG: return '(%s.nz())' % self.rv_big(e, w, s)
        return '((%s) != 0)' % self.rv(e, w, s)

    def rv_big(self, e, W, S):
        """C++ expression of type Big holding e zero-extended (only unsigned / zero-extension contexts occur)"""
        k = e[0]
        w, s = self.ty(e)
        if w <= BIG and k not in ('bin', 'cond'):
            return 'Big(%s)' % self.rv(e, w, s)
        if k == 'id':
            sm = self.sym(e[1]); return sm.cname if sm.width > BIG else 'Big(%s)' % sm.cname
        if k == 'num': return 'Big(%dull)' % e[1]
        if k == 'index':
            sm, idxs, rest = self.elem_index(e); assert not rest
            return 'RD(%s,%s,%d)' % (sm.cname, self.idx_code(sm, idxs), sm.size)
        if k == 'range':
            sm = self.sym(e[1][1]); lo = self.const(e[3]) - sm.lsb
            return '%s.shr(%d).trunc(%d)' % (sm.cname, lo, w)
        if k in ('concat', 'repl'):
            parts = e[1] if k == 'concat' else [e[2]] * self.const(e[1])
            code, pos = [], w
            for p in parts:
                pw, ps = self.ty(p); pos -= pw
                code.append('%s.shl(%d)' % (self.rv_big(p, pw, ps), pos))
            return '(' + ' | '.join(code) + ')'
        if k == 'bin':
            op = e[1]
            if op in ('|', '&', '^'): return '(%s %s %s)' % (self.rv_big(e[2], W, S), op, self.rv_big(e[3], W, S))
            if op in ('<<', '>>'):
                bw, bs = self.ty(e[3]); b = self.rv(e[3], bw, False)
                return '%s.%s((int)(%s)).trunc(%d)' % (self.rv_big(e[2], W, S), 'shl' if op == '<<' else 'shr', b, W)
            if op in ('+', '-'):
                return '(%s %s %s).trunc(%d)' % (self.rv_big(e[2], W, S), op, self.rv_big(e[3], W, S), W)
        if k == 'cond': return '(%s ? %s : %s)' % (self.nz(e[1]), self.rv_big(e[2], W, S), self.rv_big(e[3], W, S))
        if k == 'call': return 'f_%s(%s)' % (e[1], ', '.join(self.call_args(e)))
        if k == 'sys': return self.rv_big(e[2], W, S)
        raise NotImplementedError('big ' + k)

    def call_args(self, e):
        f = self.funcs[e[1]]
        ins = [d for d in f[4] if d[1] == 'input']
        out = []
        for d, a in zip(ins, e[2]):
            iw = self.const(d[4][0]) - self.const(d[4][1]) + 1 if d[4] else 1
            aw, asg = self.ty(a); cw = max(iw, aw)
            out.append(self.rv_big(a, cw, asg) + ('.trunc(%d)' % iw if iw < cw else '') if cw > BIG else '((%s) & %s)' % (self.rv(a, cw, asg), self.mask(iw)))
        return out

    # ---- assignments ----
    def lhs_width(self, l):
        if l[0] == 'concat': return sum(self.lhs_width(x) for x in l[1])
        return self.ty(l)[0]

    def assign(self, l, r, nba, out, ind):
        lw = self.lhs_width(l); rw, rs = self.ty(r); W = max(lw, rw)
        t = 't%d' % self.tmp; self.tmp += 1
        if W > BIG:
            out.append('%s{ Big %s = %s;' % (ind, t, self.rv_big(r, W, rs)))
            self.store(l, t, W, nba, out, ind + '  ', big=True)
        else:
            out.append('%s{ u64 %s = %s;' % (ind, t, self.rv(r, W, rs)))
            self.store(l, t, W, nba, out, ind + '  ', big=False)
        out.append(ind + '}')

    def store(self, l, val, W, nba, out, ind, big):
        """store the low lhs_width bits of val (u64 or Big variable name) into l"""
        if l[0] == 'concat':
            pos = self.lhs_width(l)
            for p in l[1]:
                pw = self.lhs_width(p); pos -= pw
                if big: piece = '%s.shr(%d).trunc(%d)' % (val, pos, pw) if pw > BIG else '%s.slice(%d,%d)' % (val, pos, pw)
                else: piece = '((%s >> %d) & %s)' % (val, pos, self.mask(pw))
                t = 't%d' % self.tmp; self.tmp += 1
                out.append('%s{ %s %s = %s;' % (ind, 'Big' if pw > BIG else 'u64', t, piece))
                self.store(p, t, pw, nba, out, ind + '  ', big=pw > BIG)
                out.append(ind + '}')
            return
        if l[0] == 'id': s, idxs, sel = self.sym(l[1]), [], None
        elif l[0] == 'index':
            s, idxs, rest = self.elem_index(l); sel = ('bit', rest[0]) if rest else None
        elif l[0] in ('range', 'prange'):
            inner = l[1]
            if inner[0] == 'index': s, idxs, rest = self.elem_index(inner); assert not rest
            else: s, idxs = self.sym(inner[1]), []
            sel = (l[0], l)
        else: raise ValueError(l[0])
        local = self.cur_func is not None and s.name in self.cur_func
        targets = [s.cname] if (local or not (nba or s.nba)) else (['n_' + s.cname] if nba else [s.cname, 'n_' + s.cname])
        # memories (large arrays) written by NBA use a write queue instead of a shadow copy
        queue = nba and s.size > 4096
        for tg in targets:
            if tg.startswith('n_') and s.dims and not queue:
                out.append('%sd_%s = true;' % (ind, s.cname))          # shadow array touched: commit it at the end of the clock
            if s.dims:
                ic = self.idx_code(s, idxs)
                ref = 'AT(%s,%s,%d)' % (tg, ic, s.size)
                guard = 'if (INR(%s,%d)) ' % (ic, s.size)
            else:
                ref, guard = tg, ''
            if sel is None:
                if s.width > BIG: v = val if big else 'Big(%s)' % val
                else: v = ('%s.lo64()' % val if big else val); v = '(%s & %s)' % (v, self.mask(s.width))
                if queue:
                    out.append('%sq_%s.push_back(std::make_pair((i64)%s, (u64)%s));' % (ind, s.cname, self.idx_code(s, idxs), v))
                else:
                    out.append('%s%s%s = %s;' % (ind, guard, ref, v))
            else:
                assert not queue
                if sel[0] == 'bit':
                    bw, bs = self.ty(sel[1]); b = '((%s) - %d)' % (self.rv(sel[1], bw, bs), s.lsb); w = 1
                elif sel[0] == 'range':
                    b = '%d' % (self.const(sel[1][3]) - s.lsb); w = self.const(sel[1][2]) - self.const(sel[1][3]) + 1
                else:
                    aw, asg = self.ty(sel[1][2]); a = self.rv(sel[1][2], aw, asg); w = self.const(sel[1][3])
                    b = '((%s) - %d)' % (a, s.lsb) if sel[1][4] > 0 else '((%s) - %d - %d)' % (a, s.lsb, w - 1)
                v = '%s.lo64()' % val if big else val
                if s.width > BIG:
                    out.append('%s%s%s.setslice(%s,%d,%s);' % (ind, guard, ref, b, w, v))
                else:
                    out.append('%s%s{ int sh_ = (int)(%s); %s = (%s & ~(%s << sh_)) | ((%s & %s) << sh_); }' % (ind, guard, b, ref, ref, self.mask(w), v, self.mask(w)))

    # ---- statements ----
    def stmt(self, s, out, ind):
        k = s[0]
        if k == 'block':
            for x in s[1]: self.stmt(x, out, ind)
        elif k == 'if':
            out.append('%sif (%s) {' % (ind, self.nz(s[1]))); self.stmt(s[2], out, ind + '  ')
            if s[3]:
                out.append(ind + '} else {'); self.stmt(s[3], out, ind + '  ')
            out.append(ind + '}')
        elif k == 'case':
            tys = [self.ty(s[1])] + [self.ty(l) for labels, _ in s[2] for l in labels]
            W, S = max(t[0] for t in tys), all(t[1] for t in tys)
            t = 't%d' % self.tmp; self.tmp += 1
            out.append('%s{ u64 %s = %s;' % (ind, t, self.rv(s[1], W, S)))
            first = True
            for labels, body in s[2]:
                cond = ' || '.join('%s == %s' % (t, self.rv(l, W, S)) for l in labels)
                out.append('%s%sif (%s) {' % (ind, '' if first else '} else ', cond)); first = False
                self.stmt(body, out, ind + '  ')
            if s[3]:
                out.append('%s%s{' % (ind, '} else ' if not first else '')); self.stmt(s[3], out, ind + '  ')
            out.append(ind + '}}')
        elif k == 'for':
            self.stmt(s[1], out, ind)
            out.append('%swhile (%s) {' % (ind, self.nz(s[2]))); self.stmt(s[4], out, ind + '  '); self.stmt(s[3], out, ind + '  ')
            out.append(ind + '}')
        elif k in ('ba', 'nba'):
            self.assign(s[1], s[2], k == 'nba', out, ind)
        else: raise ValueError(k)

    # ---- whole model ----
    def decl_code(self, s, prefix=''):
        ctype = 'Big' if s.width > BIG else 'u64'
        if s.dims: return 'std::vector<%s> %s%s = std::vector<%s>(%d);' % (ctype, prefix, s.cname, ctype, s.size)
        return '%s %s%s = 0;' % (ctype, prefix, s.cname)

    def generate(self):
        self.elaborate()
        o = [PRELUDE, 'struct Sim {']
        for s in self.syms.values():
            o.append('  ' + self.decl_code(s))
            if s.nba:
                if s.size > 4096: o.append('  std::vector<std::pair<i64,u64> > q_%s;' % s.cname)
                else:
                    o.append('  ' + self.decl_code(s, 'n_'))
                    if s.dims: o.append('  bool d_%s = false;' % s.cname)
        # functions
        for f in self.funcs.values():
            _, name, signed, rg, decls, body = f
            loc = {}
            for d in decls:
                _, kind, n, sg, r2, dims, init = d
                msb, lsb = (self.const(r2[0]), self.const(r2[1])) if r2 else ((31, 0) if kind == 'integer' else (0, 0))
                loc[n] = Sym(kind, n, msb - lsb + 1, sg, lsb, [])
            rw = self.const(rg[0]) - self.const(rg[1]) + 1 if rg else 1
            loc[name] = Sym('reg', name, rw, signed, self.const(rg[1]) if rg else 0, [])
            for v in loc.values(): v.cname = 'l_' + v.name
            self.cur_func = loc
            ins = [d[2] for d in decls if d[1] == 'input']
            params = ', '.join(('Big ' if loc[n].width > BIG else 'u64 ') + loc[n].cname for n in ins)
            o.append('  %s f_%s(%s) {' % ('Big' if rw > BIG else 'u64', name, params))
            for n, v in loc.items():
                if n not in ins: o.append('    %s %s = 0;' % ('Big' if v.width > BIG else 'u64', v.cname))
            body_out = []; self.stmt(body, body_out, '    '); o += body_out
            o.append('    return %s;' % loc[name].cname)
            o.append('  }')
            self.cur_func = None
        # constants + initial values
        o.append('  void init() {')
        for s in self.syms.values():
            if s.init is not None and s.kind == 'reg':
                tmp = []; self.assign(('id', s.name), s.init, False, tmp, '    '); o += tmp
        o.append('    comb();')
        for s in self.syms.values():
            if s.nba and s.size <= 4096: o.append('    n_%s = %s;' % (s.cname, s.cname))
        o.append('  }')
        # combinational: declaration assignments, assigns, always @(*)
        o.append('  void comb() {')
        for s in self.syms.values():
            if s.init is not None and s.kind == 'wire':
                tmp = []; self.assign(('id', s.name), s.init, False, tmp, '    '); o += tmp
        for it in self.items:
            if it[0] == 'assign':
                tmp = []; self.assign(it[1], it[2], False, tmp, '    '); o += tmp
            elif it[0] == 'always' and it[1] == ['*']:
                tmp = []; self.stmt(it[2], tmp, '    '); o += tmp
        o.append('  }')
        # clocked blocks
        n = 0
        for it in self.items:
            if it[0] == 'always' and it[1] != ['*']:
                assert it[1][0] == 'posedge' and it[1][1] == 'clk', it[1]
                o.append('  void blk%d() {' % n); tmp = []; self.stmt(it[2], tmp, '    '); o += tmp; o.append('  }'); n += 1
        o.append('  void clock() {    // one rising edge of clk')
        o.append('    comb();')
        for i in range(n): o.append('    blk%d();' % i)
        for s in self.syms.values():
            if s.nba:
                if s.size > 4096:
                    o.append('    for (size_t i_ = 0; i_ < q_%s.size(); i_++) if (INR(q_%s[i_].first,%d)) %s[q_%s[i_].first] = q_%s[i_].second;' % (s.cname, s.cname, s.size, s.cname, s.cname, s.cname))
                    o.append('    q_%s.clear();' % s.cname)
                elif s.dims: o.append('    if (d_%s) { %s = n_%s; d_%s = false; }' % (s.cname, s.cname, s.cname, s.cname))
                else: o.append('    %s = n_%s;' % (s.cname, s.cname))
        o.append('    comb();')
        o.append('  }')
        o.append('};')
        return '\n'.join(o) + '\n'


PRELUDE = r'''// GENERATED by oracle/vl2c.py from the reference RTL - lives only under oracle/_ref/ (git-ignored).
#include <stdint.h>
#include <string.h>
#include <utility>
#include <vector>
typedef uint64_t u64; typedef int64_t i64;
static inline u64 SX(u64 v, int w) { if (w >= 64) return v; u64 m = 1ull << (w - 1); v &= (m << 1) - 1; return (v ^ m) - m; }
static inline u64 SHL(u64 a, u64 n) { return n >= 64 ? 0 : a << n; }
static inline u64 SHR(u64 a, u64 n) { return n >= 64 ? 0 : a >> n; }
static inline u64 ASR(u64 a, int w, u64 n) { i64 v = (i64)SX(a, w); return (u64)(v >> (n >= 63 ? 63 : n)); }
static inline i64 SDIV(i64 a, i64 b) { return b ? a / b : 0; }
static inline i64 SDIVR(i64 a, i64 b) { return b ? a % b : 0; }
static inline u64 UDIV(u64 a, u64 b) { return b ? a / b : 0; }
static inline u64 UDIVR(u64 a, u64 b) { return b ? a % b : 0; }
#define INR(i, n) ((i64)(i) >= 0 && (i64)(i) < (i64)(n))
struct Big {                                    // 512-bit unsigned helper for the few wide vectors
    u64 w[8];
    Big() { memset(w, 0, sizeof w); }
    Big(u64 v) { memset(w, 0, sizeof w); w[0] = v; }
    u64 lo64() const { return w[0]; }
    bool nz() const { for (int i = 0; i < 8; i++) if (w[i]) return true; return false; }
    Big shl(int n) const { Big r; if (n >= 512 || n < 0) return r; int q = n >> 6, s = n & 63; for (int i = 7; i >= q; i--) { r.w[i] = w[i - q] << s; if (s && i - q - 1 >= 0) r.w[i] |= w[i - q - 1] >> (64 - s); } return r; }
    Big shr(int n) const { Big r; if (n >= 512 || n < 0) return r; int q = n >> 6, s = n & 63; for (int i = 0; i + q < 8; i++) { r.w[i] = w[i + q] >> s; if (s && i + q + 1 < 8) r.w[i] |= w[i + q + 1] << (64 - s); } return r; }
    Big trunc(int bits) const { Big r = *this; for (int i = 0; i < 8; i++) { int lo = i * 64; if (lo >= bits) r.w[i] = 0; else if (bits - lo < 64) r.w[i] &= (1ull << (bits - lo)) - 1; } return r; }
    u64 slice(i64 lo, int width) const { if (lo < 0 || lo >= 512) return 0; Big r = shr((int)lo); return width >= 64 ? r.w[0] : (r.w[0] & ((1ull << width) - 1)); }
    void setslice(i64 lo, int width, u64 v) { if (lo < 0 || lo + width > 512) return; Big m = Big(width >= 64 ? ~0ull : ((1ull << width) - 1)).shl((int)lo), x = Big(v).shl((int)lo); for (int i = 0; i < 8; i++) w[i] = (w[i] & ~m.w[i]) | (x.w[i] & m.w[i]); }
    Big operator|(const Big &o) const { Big r; for (int i = 0; i < 8; i++) r.w[i] = w[i] | o.w[i]; return r; }
    Big operator&(const Big &o) const { Big r; for (int i = 0; i < 8; i++) r.w[i] = w[i] & o.w[i]; return r; }
    Big operator^(const Big &o) const { Big r; for (int i = 0; i < 8; i++) r.w[i] = w[i] ^ o.w[i]; return r; }
    Big operator+(const Big &o) const { Big r; unsigned __int128 c = 0; for (int i = 0; i < 8; i++) { c += (unsigned __int128)w[i] + o.w[i]; r.w[i] = (u64)c; c >>= 64; } return r; }
    Big operator-(const Big &o) const { Big r; __int128 c = 0; for (int i = 0; i < 8; i++) { c += (__int128)w[i] - o.w[i]; r.w[i] = (u64)c; c >>= 64; } return r; }
    bool operator==(const Big &o) const { return memcmp(w, o.w, sizeof w) == 0; }
};
static inline u64 BITOF(const Big &b, i64 i) { return b.slice(i, 1); }
template <class T> static inline T RD(const std::vector<T> &v, i64 i, i64 n) { return INR(i, n) ? v[(size_t)i] : T(); }   // out-of-range read = X -> 0
#define AT(v, i, n) v[(size_t)(i)]
'''


def main():
    src, dst = sys.argv[1], sys.argv[2]
    overrides = {}
    for a in sys.argv[3:]:
        k, v = a.split('='); overrides[k] = int(v)
    items = Parser(lex(open(src).read())).module()
    g = Gen(items, overrides)
    code = g.generate()
    with open(dst, 'w') as f:
        f.write('// parameters: %s\n' % ' '.join('%s=%d' % kv for kv in sorted(overrides.items())))
        f.write(code)
    print('vl2c: %d module items, %d signals, %d functions -> %s' % (len(items), len(g.syms), len(g.funcs), dst))


if __name__ == '__main__':
    main()
