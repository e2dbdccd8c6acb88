"""f90py -- TEST INFRASTRUCTURE ONLY (never imported by the product, see oracle/README in the Makefile header).

Executes the reference's own Fortran source text without a Fortran compiler.

Neither this image nor the GPU box has gfortran / flang / nvfortran (profiles/r02_fortran_probe_*.log), so the C++
oracle could not be pinned to a binary of the reference.  This module pins it to the next best thing: the UNMODIFIED
files under /root/reference/src are read where they lie, every procedure that is called is translated statement by
statement into Python (numpy arrays for Fortran arrays, Python floats = IEEE binary64 for `real` under the reference's
-fdefault-real-8, CMakeLists.txt:27) and executed.  Expressions are evaluated exactly as written: left to right with
Fortran's precedence, no re-association, no fused multiply-add, `x**2` as `x*x`, integer division truncating,
dot_product / sum / matmul as sequential sums in index order -- i.e. what gfortran emits for x86-64 without
-ffast-math.  tests/golden/make_golden_ref.py drives solve_uvwp & co. through this and stores the outputs as the
fixtures tests/test_oracle_vs_reference_source.py compares the oracle with.

Supported subset (what src/equations/mod_uvwp.f90, src/modules/mod_solver.f90, mod_subdomains.f90, mod_eqn_setup.f90,
mod_physics.f90, mod_properties.f90 and the helpers they call in mod_util.f90 use): modules, derived types with
extension / type-bound procedures / procedure-pointer components, allocatable and pointer arrays, array sections and
whole-array expressions, do / do while / if / select case / select type / associate, subroutines with scalar
arguments passed by reference (returned as a tuple and copied back at the call site), functions, the intrinsics below.
Anything else raises F90Unsupported at translation time, naming the statement.
"""
import keyword
import math
import re

import numpy as np


class F90Unsupported(Exception):
    pass


class F90Stop(Exception):
    pass


# ------------------------------------------------------------------------------------------------ source -> statements
def _strip_comment(line):
    q = None
    for i, ch in enumerate(line):
        if q:
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
        elif ch == '!':
            return line[:i]
    return line


def read_statements(path, defines=()):
    """Logical statements of a free-form source file: comments stripped, continuations joined, ';' split, cpp
    conditionals resolved with only `defines` defined."""
    out = []
    stack = []  # cpp: [taking, taken_any]
    cur = ''
    for raw in open(path, encoding='latin-1').read().split('\n'):
        s = raw.strip()
        if s.startswith('#'):
            d = s[1:].strip()
            if d.startswith('ifdef'):
                on = d.split()[1] in defines
                stack.append([on, on])
            elif d.startswith('ifndef'):
                on = d.split()[1] not in defines
                stack.append([on, on])
            elif d.startswith('elseif') or d.startswith('elif'):
                m = re.search(r'defined\s*\(?\s*(\w+)', d)
                on = (not stack[-1][1]) and bool(m) and m.group(1) in defines
                stack[-1][0] = on
                stack[-1][1] = stack[-1][1] or on
            elif d.startswith('else'):
                stack[-1][0] = not stack[-1][1]
            elif d.startswith('endif'):
                stack.pop()
            continue
        if any(not t[0] for t in stack):
            continue
        line = _strip_comment(raw).strip()
        if not line:
            continue
        if line.startswith('&'):
            line = line[1:].lstrip()
        if line.endswith('&'):
            cur += line[:-1]
            continue
        cur += line
        # split at ';' outside strings
        parts, q, start = [], None, 0
        for i, ch in enumerate(cur):
            if q:
                if ch == q:
                    q = None
            elif ch in "'\"":
                q = ch
            elif ch == ';':
                parts.append(cur[start:i])
                start = i + 1
        parts.append(cur[start:])
        for p in parts:
            if p.strip():
                out.append(p.strip())
        cur = ''
    return out


_TOK = re.compile(r"""
    (?P<str>'(?:[^']|'')*'|"(?:[^"]|"")*")
  | (?P<dotop>\.(?:and|or|not|eqv|neqv|eq|ne|lt|le|gt|ge|true|false)\.)
  | (?P<userop>\.[a-z]+\.)
  | (?P<real>(?:\d+\.\d*(?![a-z]+\.)|\.\d+|\d+\.(?![a-z]))(?:[ed][+-]?\d+)?(?:_\w+)?|\d+[ed][+-]?\d+(?:_\w+)?)
  | (?P<int>\d+(?:_\w+)?)
  | (?P<name>[a-z_]\w*)
  | (?P<op>\*\*|//|==|/=|<=|>=|=>|\(/|/\)|::|[-+*/(),:%=<>\[\]])
  | (?P<ws>\s+)
""", re.X | re.I)


def tokenize(stmt):
    toks, pos = [], 0
    while pos < len(stmt):
        m = _TOK.match(stmt, pos)
        if not m:
            raise F90Unsupported('cannot tokenize: %r at %r' % (stmt, stmt[pos:pos + 20]))
        pos = m.end()
        k = m.lastgroup
        if k == 'ws':
            continue
        t = m.group()
        if k == 'str':
            q = t[0]
            toks.append(('str', t[1:-1].replace(q + q, q)))
        elif k == 'dotop':
            t = t.lower()
            if t in ('.true.', '.false.'):
                toks.append(('log', t == '.true.'))
            else:
                toks.append(('op', {'.eq.': '==', '.ne.': '/=', '.lt.': '<', '.le.': '<=', '.gt.': '>', '.ge.': '>='}.get(t, t)))
        elif k == 'userop':
            toks.append(('op', t.lower()))
        elif k == 'name':
            toks.append(('name', t.lower()))
        elif k == 'op' and t == '(/' and toks and toks[-1] == ('name', 'operator'):
            toks.append(('op', '('))
            toks.append(('op', '/'))
        else:
            toks.append((k, t.lower()))
    # `(/` directly followed by `=` or `)` is "( /=" resp. "( / )" -- not an array constructor; not used by the reference
    return toks


# ------------------------------------------------------------------------------------------------ expression parser
class Parser:
    def __init__(self, toks, pos=0):
        self.t, self.i = toks, pos

    def peek(self, k=0):
        return self.t[self.i + k] if self.i + k < len(self.t) else ('eof', '')

    def next(self):
        tok = self.peek()
        self.i += 1
        return tok

    def accept(self, text):
        if self.peek() == ('op', text):
            self.i += 1
            return True
        return False

    def expect(self, text):
        if not self.accept(text):
            raise F90Unsupported('expected %r at token %d of %r' % (text, self.i, self.t))

    def at_end(self):
        return self.i >= len(self.t)

    # precedence climbing
    def expr(self):
        return self.p_eqv()

    def p_eqv(self):
        a = self.p_or()
        while self.peek() in (('op', '.eqv.'), ('op', '.neqv.')):
            op = self.next()[1]
            a = ('bin', op, a, self.p_or())
        return a

    def p_or(self):
        a = self.p_and()
        while self.peek() == ('op', '.or.'):
            self.next()
            a = ('bin', '.or.', a, self.p_and())
        return a

    def p_and(self):
        a = self.p_not()
        while self.peek() == ('op', '.and.'):
            self.next()
            a = ('bin', '.and.', a, self.p_not())
        return a

    def p_not(self):
        if self.peek() == ('op', '.not.'):
            self.next()
            return ('un', '.not.', self.p_not())
        return self.p_rel()

    def p_rel(self):
        a = self.p_cat()
        if self.peek()[0] == 'op' and self.peek()[1] in ('==', '/=', '<', '<=', '>', '>='):
            op = self.next()[1]
            a = ('bin', op, a, self.p_cat())
        return a

    def p_cat(self):
        a = self.p_add()
        while self.peek() == ('op', '//'):
            self.next()
            a = ('bin', '//', a, self.p_add())
        return a

    def p_add(self):
        if self.peek() in (('op', '-'), ('op', '+')):
            op = self.next()[1]
            a = ('un', op, self.p_mul())
        else:
            a = self.p_mul()
        while self.peek() in (('op', '-'), ('op', '+')):
            op = self.next()[1]
            a = ('bin', op, a, self.p_mul())
        return a

    def p_mul(self):
        a = self.p_pow()
        while self.peek() in (('op', '*'), ('op', '/')):
            op = self.next()[1]
            a = ('bin', op, a, self.p_pow())
        return a

    def p_pow(self):
        a = self.p_primary()
        if self.peek() == ('op', '**'):
            self.next()
            if self.peek() in (('op', '-'), ('op', '+')):
                op = self.next()[1]
                b = ('un', op, self.p_pow())
            else:
                b = self.p_pow()
            a = ('bin', '**', a, b)
        return a

    def p_args(self):
        """after '(' : list of argument nodes up to ')'"""
        args = []
        if self.accept(')'):
            return args
        while True:
            args.append(self.p_arg())
            if self.accept(')'):
                return args
            self.expect(',')

    def p_arg(self):
        # keyword argument
        if self.peek()[0] == 'name' and self.peek(1) == ('op', '='):
            name = self.next()[1]
            self.next()
            return ('kw', name, self.expr())
        lo = hi = st = None
        if self.peek() != ('op', ':'):
            lo = self.expr()
            if self.peek() != ('op', ':'):
                return lo
        self.expect(':')
        if self.peek() not in (('op', ','), ('op', ')'), ('op', ':')):
            hi = self.expr()
        if self.accept(':'):
            st = self.expr()
        return ('slice', lo, hi, st)

    def p_primary(self):
        k, v = self.next()
        if k == 'int':
            return ('num', v.split('_')[0], 'int')
        if k == 'real':
            return ('num', v, 'real')
        if k == 'str':
            return ('str', v)
        if k == 'log':
            return ('log', v)
        if (k, v) == ('op', '('):
            e = self.expr()
            self.expect(')')
            return ('paren', e)
        if (k, v) in (('op', '['), ('op', '(/')):
            close = ']' if v == '[' else '/)'
            items = []
            while True:
                items.append(self.expr())
                if self.accept(close):
                    break
                self.expect(',')
            return ('arr', items)
        if k == 'name':
            segs = []
            name = v
            while True:
                args = None
                if self.accept('('):
                    args = self.p_args()
                    if self.peek() == ('op', '('):  # a(i)(1:2): substring of an array element
                        self.next()
                        segs.append((name, args))
                        name, args = None, self.p_args()
                segs.append((name, args))
                if self.accept('%'):
                    name = self.next()[1]
                    continue
                break
            return ('des', segs)
        raise F90Unsupported('unexpected token %r in %r' % ((k, v), self.t))


# ------------------------------------------------------------------------------------------------ program model
class Var:
    def __init__(self, name, base, tname=None, rank=0, dims=None, alloc=False, pointer=False, param=False, init=None, optional=False,
                 intent=None, charlen=None, iface=None):
        self.name, self.base, self.tname, self.rank, self.dims = name, base, tname, rank, dims
        self.alloc, self.pointer, self.param, self.init, self.optional, self.intent = alloc, pointer, param, init, optional, intent
        self.charlen, self.iface = charlen, iface
        self.dummy = False

    def T(self):
        return (self.base, self.tname, self.rank)


class TypeDef:
    def __init__(self, name, parent):
        self.name, self.parent, self.comps, self.bound = name, parent, {}, {}
        self.order = []


class Proc:
    def __init__(self, name, kind, args, result, stmts, module, prefix_type=None):
        self.name, self.kind, self.args, self.result, self.stmts, self.module = name, kind, args, result, stmts, module
        self.prefix_type = prefix_type
        self.vars = None   # name -> Var after analysis
        self.outs = None   # dummy names returned to the caller
        self.body_start = 0


_TYPE_KW = ('integer', 'real', 'logical', 'character', 'type', 'class', 'procedure', 'double', 'complex')
_PY_RESERVED = set(keyword.kwlist) | {'np', 'math', 'max', 'min', 'abs', 'int', 'float', 'len', 'sum', 'range', 'print', 'str', 'bool', 'list', 'tuple',
                                      'type', 'object', 'id', 'all', 'any', 'round', 'pow', 'map', 'filter', 'iter', 'next', 'set', 'dict', 'input', 'copy'}


def split_top(toks, sep=','):
    """split a token list at top-level separators"""
    parts, depth, cur = [], 0, []
    for tk in toks:
        if tk[0] == 'op' and tk[1] in ('(', '[', '(/'):
            depth += 1
        elif tk[0] == 'op' and tk[1] in (')', ']', '/)'):
            depth -= 1
        if depth == 0 and tk == ('op', sep):
            parts.append(cur)
            cur = []
        else:
            cur.append(tk)
    parts.append(cur)
    return parts


def match_paren(toks, i):
    """toks[i] == '(' -> index of the matching ')'"""
    depth = 0
    for j in range(i, len(toks)):
        if toks[j][0] == 'op' and toks[j][1] in ('(', '(/'):
            depth += 1
        elif toks[j][0] == 'op' and toks[j][1] in (')', '/)'):
            depth -= 1
            if depth == 0:
                return j
    raise F90Unsupported('unbalanced parentheses in %r' % (toks,))


def parse_decl(toks):
    """A type declaration statement -> list of Var, or None if `toks` is not one."""
    if not toks or toks[0][0] != 'name' or toks[0][1] not in _TYPE_KW:
        return None
    kw = toks[0][1]
    i = 1
    base, tname, charlen, iface = None, None, None, None
    if kw == 'double':
        base = 'real'
        i = 2
    elif kw in ('integer', 'real', 'logical', 'complex'):
        base = {'integer': 'int', 'real': 'real', 'logical': 'logical', 'complex': 'complex'}[kw]
        if i < len(toks) and toks[i] == ('op', '('):
            i = match_paren(toks, i) + 1
        elif i < len(toks) and toks[i] == ('op', '*'):
            i += 2
    elif kw == 'character':
        base = 'char'
        if i < len(toks) and toks[i] == ('op', '('):
            j = match_paren(toks, i)
            charlen = toks[i + 1:j]
            i = j + 1
        elif i < len(toks) and toks[i] == ('op', '*'):
            i += 2
    elif kw in ('type', 'class'):
        if i >= len(toks) or toks[i] != ('op', '('):
            return None  # a type definition, not a declaration
        j = match_paren(toks, i)
        base, tname = 'type', toks[i + 1][1]
        i = j + 1
    elif kw == 'procedure':
        base = 'proc'
        if i < len(toks) and toks[i] == ('op', '('):
            j = match_paren(toks, i)
            iface = toks[i + 1][1] if j > i + 1 else None
            i = j + 1
    # a function statement such as `integer function sgn(x)` is not a declaration
    if any(t == ('name', 'function') for t in toks[i:i + 4]) and ('op', '::') not in toks:
        return None
    attrs = {}
    dims = None
    if ('op', '::') in toks:
        k = toks.index(('op', '::'))
        for a in split_top(toks[i:k]):
            if not a:
                continue
            an = a[0][1]
            if an == 'dimension':
                dims = a[2:match_paren(a, 1)]
            elif an == 'intent':
                attrs['intent'] = ''.join(t[1] for t in a[2:-1])
            else:
                attrs[an] = True
        ents = toks[k + 1:]
    else:
        ents = toks[i:]
    out = []
    for e in split_top(ents):
        if not e:
            continue
        name = e[0][1]
        j = 1
        edims = dims
        if j < len(e) and e[j] == ('op', '('):
            m = match_paren(e, j)
            edims = e[j + 1:m]
            j = m + 1
        if j < len(e) and e[j] == ('op', '*'):  # character*len
            j += 2
        init = None
        if j < len(e) and e[j] in (('op', '='), ('op', '=>')):
            init = e[j + 1:]
        dlist, rank = None, 0
        if edims is not None:
            dlist = []
            for d in split_top(edims):
                if ('op', ':') in d:
                    c = split_top(d, ':')
                    dlist.append((c[0] or None, c[1] or None))
                elif d == [('op', '*')]:
                    dlist.append((None, None))
                else:
                    dlist.append((None, d))
            rank = len(dlist)
        out.append(Var(name, base, tname, rank, dlist, alloc='allocatable' in attrs, pointer='pointer' in attrs, param='parameter' in attrs,
                       init=init, optional='optional' in attrs, intent=attrs.get('intent'), charlen=charlen, iface=iface))
    return out


class World:
    """All parsed modules; translated procedures live in self.ns (one Python namespace)."""

    def __init__(self, files, defines=()):
        self.types, self.procs, self.modvars, self.generics = {}, {}, {}, {}
        self.byname = {}  # procedure name -> keys in self.procs (the same name may be defined in several modules)
        self.ns = dict(_RUNTIME)
        self.ns['_records'] = []
        self.ns['_world'] = self
        self.translated = set()
        self.pending_init = []
        for f in files:
            self._parse_file(f, defines)
        for td in self.types.values():
            self._emit_type(td)
        self._init_modvars()

    # ---- parsing into units
    def _parse_file(self, path, defines):
        stmts = [tokenize(s) for s in read_statements(path, defines)]
        i, module = 0, None
        n = len(stmts)

        def is_end(t, what):
            if t[0] != ('name', 'end') and not (t[0][0] == 'name' and t[0][1] == 'end' + what):
                return False
            if t[0][1] == 'end' + what:
                return True
            return len(t) == 1 or t[1] == ('name', what)

        while i < n:
            t = stmts[i]
            if t[0] == ('name', 'module') and len(t) == 2:
                module = t[1][1]
                i += 1
                continue
            if is_end(t, 'module') or t[0] == ('name', 'contains') or is_end(t, 'program'):
                i += 1
                continue
            if t[0] == ('name', 'program'):
                module = '__main__'
                # skip the main program's body
                while i < n and not is_end(stmts[i], 'program'):
                    i += 1
                continue
            # interface blocks: record generic -> specific names, skip the rest
            if t[0] == ('name', 'interface') or t[:2] == [('name', 'abstract'), ('name', 'interface')]:
                gname = t[1][1] if len(t) > 1 and t[1][0] == 'name' and t[0][1] == 'interface' else None
                i += 1
                while not is_end(stmts[i], 'interface'):
                    s = stmts[i]
                    if gname and s[:2] == [('name', 'module'), ('name', 'procedure')]:
                        self.generics.setdefault(gname, []).extend(x[1] for x in s[2:] if x[0] == 'name')
                    elif gname and s[0] == ('name', 'procedure') and len(s) > 1:
                        self.generics.setdefault(gname, []).extend(x[1] for x in s[1:] if x[0] == 'name')
                    i += 1
                i += 1
                continue
            # type definition
            if t[0] == ('name', 'type') and (len(t) < 2 or t[1] != ('op', '(')):
                names = [x for x in t if x[0] == 'name']
                parent = None
                if ('name', 'extends') in t:
                    parent = t[t.index(('name', 'extends')) + 2][1]
                td = TypeDef(names[-1][1], parent)
                i += 1
                in_bound = False
                while not is_end(stmts[i], 'type'):
                    s = stmts[i]
                    i += 1
                    if s[0] == ('name', 'contains'):
                        in_bound = True
                        continue
                    if s[0][1] in ('private', 'public', 'sequence'):
                        continue
                    if in_bound:
                        if s[0] == ('name', 'procedure'):
                            k = s.index(('op', '::')) if ('op', '::') in s else 0
                            for e in split_top(s[k + 1:]):
                                if len(e) >= 3 and e[1] == ('op', '=>'):
                                    td.bound[e[0][1]] = e[2][1]
                                elif e:
                                    td.bound[e[0][1]] = e[0][1]
                        continue
                    d = parse_decl(s)
                    if d is None:
                        raise F90Unsupported('in type %s: %r' % (td.name, s))
                    for v in d:
                        td.comps[v.name] = v
                        td.order.append(v.name)
                i += 1
                self.types[td.name] = td
                continue
            # procedures
            hdr = self._proc_header(t)
            if hdr:
                kind, name, args, result, ptype = hdr
                j = i + 1
                depth = 1
                body = []
                while True:
                    s = stmts[j]
                    if self._proc_header(s) and not (s[0] == ('name', 'end')):
                        depth += 1
                    if is_end(s, 'subroutine') or is_end(s, 'function') or (s == [('name', 'end')]):
                        depth -= 1
                        if depth == 0:
                            break
                    body.append(s)
                    j += 1
                key = name if name not in self.byname else '%s__%s' % (name, module)
                if name in self.byname and len(self.byname[name]) == 1:  # the first definition gets its module suffix too
                    k0 = self.byname[name][0]
                    if '__' not in k0:
                        p0 = self.procs.pop(k0)
                        k1 = '%s__%s' % (name, p0.module)
                        p0.key = k1
                        self.procs[k1] = p0
                        self.byname[name][0] = k1
                pr = Proc(name, kind, args, result, body, module, ptype)
                pr.key = key
                self.procs[key] = pr
                self.byname.setdefault(name, []).append(key)
                i = j + 1
                continue
            # module-level declarations, old-style `parameter (name = value)` and `data name / values /` statements
            if module and module != '__main__':
                d = parse_decl(t)
                if d:
                    for v in d:
                        v.module = module
                        self.modvars[v.name] = v
                elif t[0] == ('name', 'parameter') and t[1] == ('op', '('):
                    for item in split_top(t[2:match_paren(t, 1)]):
                        if item[0][1] in self.modvars and item[1] == ('op', '='):
                            self.modvars[item[0][1]].param = True
                            self.modvars[item[0][1]].init = item[2:]
                elif t[0] == ('name', 'data') and len(t) > 3 and t[1][0] == 'name' and t[2] == ('op', '/') and t[-1] == ('op', '/'):
                    if t[1][1] in self.modvars:
                        self.modvars[t[1][1]].data = [x for x in split_top(t[3:-1])]
            i += 1

    @staticmethod
    def _proc_header(t):
        names = [x[1] for x in t[:8] if x[0] == 'name']
        for kind in ('subroutine', 'function'):
            if ('name', kind) in t[:8]:
                k = t.index(('name', kind))
                if t[0] == ('name', 'end'):
                    return None
                pre = t[:k]
                ok_pre = all(x[0] == 'name' and x[1] in ('pure', 'elemental', 'recursive', 'impure', 'integer', 'real', 'logical', 'double', 'precision',
                                                         'character') or x[0] in ('op', 'int') for x in pre)
                if not ok_pre or k + 1 >= len(t) or t[k + 1][0] != 'name':
                    return None
                name = t[k + 1][1]
                args, result = [], None
                if k + 2 < len(t) and t[k + 2] == ('op', '('):
                    m = match_paren(t, k + 2)
                    args = [x[1] for x in t[k + 3:m] if x[0] == 'name']
                    rest = t[m + 1:]
                    if rest and rest[0] == ('name', 'result'):
                        result = rest[2][1]
                ptype = None
                for x in pre:
                    if x[0] == 'name' and x[1] in ('integer', 'real', 'logical', 'double', 'character'):
                        ptype = {'integer': 'int', 'real': 'real', 'logical': 'logical', 'double': 'real', 'character': 'char'}[x[1]]
                return kind, name, args, result, ptype
        return None

    # ---- names
    def resolve(self, name, module=None):
        """key in self.procs of procedure `name` as seen from `module` (its own definition first), or None"""
        keys = self.byname.get(name)
        if not keys:
            return None
        if len(keys) == 1:
            return keys[0]
        for k in keys:
            if self.procs[k].module == module:
                return k
        return keys[-1]

    def pyname(self, n):
        if n in _PY_RESERVED or n in self.byname or n in _INTRINSICS or n.startswith('_'):
            return n + '_v'
        return n

    @staticmethod
    def attr(n):
        return n + '_' if keyword.iskeyword(n) else n

    def extends(self, tname, ancestor):
        """is type `tname` the type `ancestor` or an extension of it"""
        while tname is not None:
            if tname == ancestor:
                return True
            td = self.types.get(tname)
            tname = td.parent if td else None
        return False

    def comp(self, tname, cname):
        td = self.types.get(tname)
        while td:
            if cname in td.comps:
                return td.comps[cname]
            td = self.types.get(td.parent) if td.parent else None
        return None

    def bound(self, tname, pname):
        td = self.types.get(tname)
        while td:
            if pname in td.bound:
                return td.bound[pname]
            td = self.types.get(td.parent) if td.parent else None
        return None

    # ---- derived types -> Python classes
    def _emit_type(self, td):
        if 'T_' + td.name in self.ns:
            return
        if td.parent and 'T_' + td.parent not in self.ns:
            self._emit_type(self.types[td.parent])
        world = self
        parent = self.ns['T_' + td.parent] if td.parent else object

        def init(obj):
            if td.parent:
                parent.__init__(obj)
            for cn in td.order:
                v = td.comps[cn]
                setattr(obj, world.attr(cn), world.default_value(v, None))

        self.ns['T_' + td.name] = type('T_' + td.name, (parent,), {'__init__': init})

    def default_value(self, v, cg):
        """initial value of a variable / component (cg: a code generator whose scope evaluates dimension expressions)"""
        if v.alloc or v.pointer or v.base == 'proc':
            return None
        if v.rank == 0:
            if v.init is not None:
                return self.const_eval(v.init)
            if v.base == 'type':
                return self.ns['T_' + v.tname]()
            return {'int': 0, 'real': 0.0, 'logical': False, 'char': '', 'complex': 0j}[v.base]
        shape = tuple(int(self.const_eval(hi)) - (int(self.const_eval(lo)) if lo else 1) + 1 for lo, hi in v.dims)
        arr = _newarr(shape, v.base, self.ns.get('T_' + v.tname) if v.base == 'type' else None)
        if v.init is not None:
            arr[...] = self.const_eval(v.init)
        return arr

    def const_eval(self, toks):
        cg = CodeGen(self, None)
        code, _ = cg.gen(Parser(list(toks)).expr())
        return eval(code, self.ns)

    def _init_modvars(self):
        # parameters first (they may size the others), in file order
        for v in self.modvars.values():
            if v.param and v.init is not None:
                try:
                    self.ns[self.pyname(v.name)] = self.const_eval(v.init) if v.rank == 0 else np.array(self.const_eval(v.init))
                except Exception:
                    pass
        for v in self.modvars.values():
            if not v.param:
                try:
                    val = self.default_value(v, None)
                    if getattr(v, 'data', None) is not None:  # data statement: array element order
                        vals = [self.const_eval(x) for x in v.data]
                        if v.rank == 0:
                            val = vals[0]
                        else:
                            flat = val.reshape(-1, order='F')
                            flat[:len(vals)] = vals
                            val = flat.reshape(val.shape, order='F')
                    self.ns[self.pyname(v.name)] = val
                except Exception:
                    self.ns[self.pyname(v.name)] = None

    # ---- procedures
    def analyse(self, name):
        p = self.procs[name]
        if p.vars is not None:
            return p
        p.vars = {}
        i = 0
        for i, s in enumerate(p.stmts):
            d = parse_decl(s)
            if d:
                for v in d:
                    if v.name in p.vars:  # attribute statements / redeclaration: merge
                        continue
                    p.vars[v.name] = v
                continue
            h = s[0][1] if s[0][0] == 'name' else ''
            if h in ('use', 'implicit', 'save', 'private', 'public', 'external', 'intrinsic', 'import'):
                continue
            if h == 'interface':
                continue
            break
        else:
            i = len(p.stmts)
        # skip interface blocks inside the specification part
        p.body_start = i
        for a in p.args:
            if a in p.vars:
                p.vars[a].dummy = True
            else:  # dummy procedure or implicitly typed: treat as untyped
                p.vars[a] = Var(a, 'unknown')
                p.vars[a].dummy = True
        if p.kind == 'function':
            r = p.result or p.name
            if r not in p.vars:
                p.vars[r] = Var(r, p.prefix_type or 'real')
        # dummies handed back to the caller: scalars that may be assigned, allocatable/pointer dummies that may be (re)bound
        outs = []
        assigned, called, rebound = set(), set(), set()
        for s in p.stmts[p.body_start:]:
            self._scan_writes(s, assigned, called, rebound)
        for a in p.args:
            v = p.vars[a]
            if v.intent == 'in' or v.base in ('proc',):
                continue
            scalar = v.rank == 0 and v.base in ('int', 'real', 'logical', 'char', 'unknown')
            if scalar and (a in assigned or a in called):
                outs.append(a)
            elif (v.alloc or v.pointer) and (a in rebound or (a in assigned)):
                outs.append(a)
        p.outs = outs
        return p

    def _scan_writes(self, s, assigned, called, rebound):
        # one-line if: scan the trailing statement
        if s[0] == ('name', 'if') and len(s) > 1 and s[1] == ('op', '('):
            m = match_paren(s, 1)
            rest = s[m + 1:]
            if rest and rest != [('name', 'then')]:
                self._scan_writes(rest, assigned, called, rebound)
            return
        h = s[0][1] if s[0][0] == 'name' else ''
        if h == 'call':
            for tk in s[2:]:
                if tk[0] == 'name':
                    called.add(tk[1])
            return
        if h in ('allocate', 'deallocate', 'nullify'):
            for part in split_top(s[2:match_paren(s, 1)]):
                if part and part[0][0] == 'name' and len([x for x in part if x == ('op', '%')]) == 0:
                    rebound.add(part[0][1])
            return
        if h == 'do' and len(s) > 2 and s[2] == ('op', '='):
            assigned.add(s[1][1])
            return
        if h == 'read':
            for tk in s[1:]:
                if tk[0] == 'name':
                    assigned.add(tk[1])
            return
        # assignment: name [ (..) ] = / =>   (top-level '=')
        depth = 0
        for k, tk in enumerate(s):
            if tk[0] == 'op' and tk[1] in ('(', '(/', '['):
                depth += 1
            elif tk[0] == 'op' and tk[1] in (')', '/)', ']'):
                depth -= 1
            elif depth == 0 and tk in (('op', '='), ('op', '=>')):
                if s[0][0] == 'name' and ('op', '%') not in s[:k]:
                    if tk[1] == '=>':
                        rebound.add(s[0][1])
                    else:
                        assigned.add(s[0][1])
                break

    def get(self, name, module=None):
        """the Python function of procedure `name` (as seen from `module`), translated on first use"""
        key = self.resolve(name.lower(), module)
        if key is None:
            raise F90Unsupported('procedure %s not found in the parsed sources' % name)
        if key not in self.translated:
            self.translate(key)
        return self.ns[key]

    def translate(self, name):
        if name in self.translated:
            return
        if name not in self.procs:
            raise F90Unsupported('procedure %s not found in the parsed sources' % name)
        self.translated.add(name)
        p = self.analyse(name)
        cg = CodeGen(self, p)
        src = cg.procedure()
        p.pysrc = src
        try:
            exec(compile(src, '<f90:%s>' % name, 'exec'), self.ns)
        except SyntaxError as ex:
            raise F90Unsupported('generated code for %s does not compile: %s\n%s' % (name, ex, src))
        for dep in cg.deps:
            try:
                self.translate(dep)
            except F90Unsupported as ex:  # only a problem if the procedure is actually called
                def stub(*a, _msg='%s: %s' % (dep, ex), **k):
                    raise F90Unsupported('called a procedure that could not be translated -- ' + _msg)
                self.ns[dep] = stub


# ------------------------------------------------------------------------------------------------ code generation
def _arith_T(a, b):
    base = 'real' if 'real' in (a[0], b[0]) else ('int' if (a[0], b[0]) == ('int', 'int') else ('unknown' if 'unknown' in (a[0], b[0]) else a[0]))
    return (base, None, max(a[2], b[2]))


class CodeGen:
    def __init__(self, world, proc):
        self.w, self.p = world, proc
        self.vars = dict(proc.vars) if proc else {}
        self.lines, self.ind = [], 1
        self.deps = set()
        self.globals_written = set()
        self.tmp = 0
        self.blocks = []

    # ---- helpers
    def emit(self, s):
        self.lines.append('    ' * self.ind + s)

    def lookup(self, name):
        if name in self.vars:
            return self.vars[name], False
        if name in self.w.modvars:
            return self.w.modvars[name], True
        return None, False

    def newtmp(self):
        self.tmp += 1
        return '_t%d' % self.tmp

    def rproc(self, name):
        return self.w.resolve(name, self.p.module if self.p else None)

    # ---- expressions
    def gen(self, n):
        k = n[0]
        if k == 'num':
            if n[2] == 'int':
                return str(int(n[1])), ('int', None, 0)
            t = n[1].split('_')[0].replace('d', 'e')
            return repr(float(t)), ('real', None, 0)
        if k == 'str':
            return repr(n[1].rstrip()), ('char', None, 0)
        if k == 'log':
            return ('True' if n[1] else 'False'), ('logical', None, 0)
        if k == 'paren':
            c, T = self.gen(n[1])
            return '(' + c + ')', T
        if k == 'arr':
            parts = [self.gen(x) for x in n[1]]
            base = 'real' if any(T[0] == 'real' for _, T in parts) else parts[0][1][0]
            if any(T[2] > 0 for _, T in parts):
                return '_cat([%s])' % ', '.join(c for c, _ in parts), (base, None, 1)
            dt = {'real': 'float', 'int': 'np.int64', 'logical': 'bool'}.get(base, 'object')
            return 'np.array([%s], dtype=%s)' % (', '.join(c for c, _ in parts), dt), (base, None, 1)
        if k == 'un':
            c, T = self.gen(n[2])
            if n[1] == '.not.':
                return '(not %s)' % c, ('logical', None, T[2])
            return '(%s%s)' % (n[1], c), T
        if k == 'bin':
            op = n[1]
            a, Ta = self.gen(n[2])
            b, Tb = self.gen(n[3])
            if op in ('.and.', '.or.'):
                return '(%s %s %s)' % (a, op.strip('.'), b), ('logical', None, max(Ta[2], Tb[2]))
            if op in ('.eqv.', '.neqv.'):
                return '(bool(%s) %s bool(%s))' % (a, '==' if op == '.eqv.' else '!=', b), ('logical', None, 0)
            if op in ('==', '/=', '<', '<=', '>', '>='):
                pyop = '!=' if op == '/=' else op
                if Ta[0] == 'char' or Tb[0] == 'char':
                    return '(_s(%s) %s _s(%s))' % (a, pyop, b), ('logical', None, 0)
                return '(%s %s %s)' % (a, pyop, b), ('logical', None, max(Ta[2], Tb[2]))
            if op == '//':
                return '(%s + %s)' % (a, b), ('char', None, 0)
            T = _arith_T(Ta, Tb)
            if op == '/':
                if Ta[0] == 'int' and Tb[0] == 'int':
                    return '_idiv(%s, %s)' % (a, b), T
                if 'real' in (Ta[0], Tb[0]):
                    return '(%s / %s)' % (a, b), T
                return '_div(%s, %s)' % (a, b), T
            if op == '**':
                if n[3][0] == 'num' and n[3][2] == 'int':
                    e = int(n[3][1])
                    if e == 2:
                        return '_sq(%s)' % a, Ta
                    return '_powi(%s, %d)' % (a, e), Ta
                if Tb[0] == 'int':
                    return '_powi(%s, %s)' % (a, b), Ta
                return '_powr(%s, %s)' % (a, b), ('real', None, max(Ta[2], Tb[2]))
            return '(%s %s %s)' % (a, op, b), T
        if k == 'des':
            return self.gen_des(n[1])
        if k == 'kw':
            c, T = self.gen(n[2])
            return '%s=%s' % (self.w.pyname(n[1]), c), T
        raise F90Unsupported('expression node %r' % (n,))

    def index_code(self, args, v):
        """subscripts of variable/component v -> ('[...]', rank of the result)"""
        dims = v.dims or [(None, None)] * len(args)
        if len(args) != v.rank and not (v.rank == 1 and len(args) == 1):
            if len(args) != len(dims):
                raise F90Unsupported('%d subscripts for rank-%d %s' % (len(args), v.rank, v.name))
        subs, rank = [], 0
        for a, d in zip(args, dims):
            lo = d[0] if d else None
            lb = '1'
            if lo:
                lb, _ = self.gen(Parser(list(lo)).expr())
            if a[0] == 'slice':
                rank += 1
                if a[3] is not None:
                    st, _ = self.gen(a[3])
                    l = self._shift(a[1], lb) if a[1] is not None else ''
                    h = self._shift(a[2], lb, 1) if a[2] is not None else ''
                    subs.append('_sl(%s, %s, %s, %s)' % (self.gen(a[1])[0] if a[1] is not None else 'None',
                                                         self.gen(a[2])[0] if a[2] is not None else 'None', st, lb))
                    continue
                l = self._shift(a[1], lb) if a[1] is not None else ''
                h = self._shift(a[2], lb, 1) if a[2] is not None else ''
                subs.append('%s:%s' % (l, h))
            else:
                c, T = self.gen(a)
                if T[2] > 0:  # vector subscript
                    rank += 1
                    subs.append('(%s) - %s' % (c, lb))
                else:
                    subs.append(self._shift(a, lb))
        return '[' + ', '.join(subs) + ']', rank

    def _shift(self, node, lb, plus=0):
        c, _ = self.gen(node)
        off = plus - (int(lb) if re.fullmatch(r'-?\d+', lb) else 0)
        sym = '' if re.fullmatch(r'-?\d+', lb) else ' - (%s)' % lb
        if re.fullmatch(r'-?\d+', c):
            return str(int(c) + off) + sym
        if off == 0:
            return c + sym
        return '%s %s %d%s' % (c, '+' if off > 0 else '-', abs(off), sym)

    def gen_des(self, segs, target=False):
        name, args = segs[0]
        v, is_global = self.lookup(name)
        rest = segs[1:]
        if v is None:
            if len(segs) == 1 or True:
                # function reference (intrinsic, module procedure, generic) -- never a target
                if args is None and not rest:
                    key = self.rproc(name)
                    if key is None:
                        raise F90Unsupported('unknown name %s' % name)
                    self.deps.add(key)  # a procedure name as a value (procedure pointer target, actual argument)
                    return key, ('proc', None, 0)
                code, T = self.gen_call_expr(name, args or [])
                if rest:
                    return self.gen_tail(code, T, rest, target)
                return code, T
        code = self.w.pyname(name)
        if is_global and target:
            self.globals_written.add(code)
        T = v.T()
        cur = v
        if v.base == 'unknown' and args is not None:  # dummy procedure
            cargs = ', '.join(self.gen(a)[0] for a in args)
            return '%s(%s)' % (code, cargs), ('unknown', None, 0)
        if args is not None:
            if v.rank > 0:
                ic, r = self.index_code(args, v)
                code += ic
                T = (v.base, v.tname, r)
            elif v.base == 'char':
                code += self.substr(args)
            else:
                raise F90Unsupported('subscript on scalar %s' % name)
        return self.gen_tail(code, T, rest, target)

    def substr(self, args):
        a = args[0]
        if a[0] != 'slice':
            raise F90Unsupported('substring')
        lo = self._shift(a[1], '1') if a[1] is not None else ''
        hi = self.gen(a[2])[0] if a[2] is not None else ''
        return '[%s:%s]' % (lo, hi)

    def gen_tail(self, code, T, rest, target):
        for (cname, args) in rest:
            if cname is None:  # substring of an element
                code += self.substr(args)
                continue
            if T[0] != 'type':
                raise F90Unsupported('component %s of a non-derived value %s' % (cname, code))
            comp = self.w.comp(T[1], cname)
            if comp is None:
                bp = self.w.bound(T[1], cname)
                bp = self.rproc(bp) if bp else None
                if bp is None:
                    raise F90Unsupported('type %s has no component %s' % (T[1], cname))
                self.deps.add(bp)
                p = self.w.analyse(bp)
                rT = p.vars[p.result or p.name].T() if p.kind == 'function' else ('unknown', None, 0)
                if p.kind == 'function':
                    code, T = self.fn_call(bp, p, [code] + [self.gen(a)[0] for a in (args or [])], [None] + list(args or [])), rT
                else:
                    code, T = '%s(%s)' % (bp, ', '.join([code] + [self.gen(a)[0] for a in (args or [])])), rT
                continue
            if T[2] > 0:
                raise F90Unsupported('component of an array section: %s%%%s' % (code, cname))
            if comp.base == 'proc':
                if args is None:
                    code, T = '%s.%s' % (code, self.w.attr(cname)), ('proc', None, 0)
                else:
                    cargs = ', '.join([code] + [self.gen(a)[0] for a in args])
                    code, T = '%s.%s(%s)' % (code, self.w.attr(cname), cargs), ('unknown', None, 0)
                continue
            code = '%s.%s' % (code, self.w.attr(cname))
            T = comp.T()
            if args is not None:
                if comp.rank > 0:
                    ic, r = self.index_code(args, comp)
                    code += ic
                    T = (comp.base, comp.tname, r)
                elif comp.base == 'char':
                    code += self.substr(args)
                else:
                    raise F90Unsupported('subscript on scalar component %s' % cname)
        return code, T

    def fn_call(self, key, p, codes, args):
        """a reference to function `key` whose actual arguments have the codes `codes` (nodes `args`, None for the passed
        object of a type-bound call): plain call, or -- when the function writes by-reference dummies -- an expression
        that calls it, stores the returned dummies into the actual arguments and yields the result"""
        callc = '%s(%s)' % (key, ', '.join(codes))
        if not p.outs:
            return callc
        t = self.newtmp()
        parts = ['(%s := %s)' % (t, callc)]
        for i, o in enumerate(p.outs):
            pos = p.args.index(o)
            a = None
            if pos < len(args) and (args[pos] is None or args[pos][0] != 'kw'):
                a = args[pos]
            else:
                for x in args:
                    if x is not None and x[0] == 'kw' and x[1] == o:
                        a = x[2]
            if a is None or a[0] != 'des':
                continue
            try:
                v0, _ = self.lookup(a[1][0][0])
                if v0 is None:
                    continue
                tgt = self.gen_des(a[1], target=True)[0]
            except F90Unsupported:
                continue
            val = '%s[%d]' % (t, i + 1)
            if tgt.isidentifier():
                parts.append('(%s := %s)' % (tgt, val))
            elif tgt.endswith(']'):
                depth, k = 0, len(tgt) - 1
                while k >= 0:
                    if tgt[k] == ']':
                        depth += 1
                    elif tgt[k] == '[':
                        depth -= 1
                        if depth == 0:
                            break
                    k -= 1
                parts.append('%s.__setitem__(%s, %s)' % (tgt[:k], tgt[k + 1:-1], val))
            else:
                obj, attr = tgt.rsplit('.', 1)
                parts.append('setattr(%s, %r, %s)' % (obj, attr, val))
        parts.append('%s[0]' % t)
        return '(' + ', '.join(parts) + ')[-1]'

    def gen_call_expr(self, name, args):
        if name in _INTRINSICS and name not in self.w.byname:
            return _INTRINSICS[name](self, args)
        target = name
        if name in self.w.generics and name not in self.w.byname:
            target = self.w.generics[name][0]
        if target in self.w.types and target not in self.w.byname:  # structure constructor
            raise F90Unsupported('structure constructor %s' % name)
        target = self.rproc(target)
        if target is None:
            raise F90Unsupported('unknown function %s' % name)
        p = self.w.analyse(target)
        self.deps.add(target)
        rv = p.vars[p.result or p.name]
        return self.fn_call(target, p, [self.gen(a)[0] for a in args], list(args)), rv.T()

    # ---- statements
    def procedure(self):
        p = self.p
        body = p.stmts[p.body_start:]
        for s in body:
            self.stmt(s)
        if self.blocks:
            raise F90Unsupported('unterminated block in %s' % p.name)
        self.emit(self.ret())
        head = []
        params = []
        for a in p.args:
            v = p.vars[a]
            params.append(self.w.pyname(a) + ('=None' if v.optional else ''))
        head.append('def %s(%s):' % (p.key, ', '.join(params)))
        if self.globals_written:
            head.append('    global ' + ', '.join(sorted(self.globals_written)))
        # locals
        pre = CodeGen(self.w, p)
        for name, v in p.vars.items():
            py = self.w.pyname(name)
            if v.dummy:
                if v.rank >= 2 and v.dims and all(d[1] is not None for d in v.dims[:-1]):
                    # explicit-shape dummy of rank >= 2: sequence association with the actual argument (a view in Fortran order)
                    ext = []
                    for lo, hi in v.dims:
                        if hi is None:
                            ext.append('-1')
                        else:
                            h = pre.gen(Parser(list(hi)).expr())[0]
                            l = pre.gen(Parser(list(lo)).expr())[0] if lo else '1'
                            ext.append('(%s) - (%s) + 1' % (h, l))
                    head.append('    %s = _reshape(%s, (%s,))' % (py, py, ', '.join(ext)))
                elif v.rank == 1 and v.dims and v.dims[0][1] is not None and not v.alloc and not v.pointer:
                    # explicit-shape dummy of rank 1: the dummy IS its declared extent (whole-array operations such as
                    # minval(x) or x = 0 must not see the rest of a longer actual argument)
                    lo, hi = v.dims[0]
                    if not (len(hi) == 1 and hi[0] == ('op', '*')):
                        h = pre.gen(Parser(list(hi)).expr())[0]
                        l = pre.gen(Parser(list(lo)).expr())[0] if lo else '1'
                        head.append('    %s = _view1(%s, (%s) - (%s) + 1)' % (py, py, h, l))
                continue
            if v.param:
                head.append('    %s = %s' % (py, pre.gen(Parser(list(v.init)).expr())[0]))
                continue
            if v.alloc or v.pointer or v.base == 'proc':
                head.append('    %s = None' % py)
            elif v.rank == 0:
                if v.init is not None:
                    head.append('    %s = %s' % (py, pre.gen(Parser(list(v.init)).expr())[0]))
                elif v.base == 'type':
                    head.append('    %s = T_%s()' % (py, v.tname))
                else:
                    head.append('    %s = %s' % (py, {'int': '0', 'real': '0.0', 'logical': 'False', 'char': "''", 'complex': '0j'}.get(v.base, 'None')))
            else:
                ext = []
                for lo, hi in v.dims:
                    if hi is None:
                        raise F90Unsupported('local array %s with deferred shape in %s' % (name, p.name))
                    h = pre.gen(Parser(list(hi)).expr())[0]
                    l = pre.gen(Parser(list(lo)).expr())[0] if lo else '1'
                    ext.append('(%s) - (%s) + 1' % (h, l))
                tcls = 'T_%s' % v.tname if v.base == 'type' else 'None'
                head.append('    %s = _newarr((%s,), %r, %s)' % (py, ', '.join(ext), v.base, tcls))
                if v.init is not None:
                    head.append('    %s[...] = %s' % (py, pre.gen(Parser(list(v.init)).expr())[0]))
        self.deps |= pre.deps
        return '\n'.join(head + self.lines) + '\n'

    def ret(self):
        p = self.p
        if p.kind == 'function':
            if p.outs:  # a function that also writes by-reference dummies: (result, dummies...)
                return 'return (' + ', '.join([self.w.pyname(p.result or p.name)] + [self.w.pyname(a) for a in p.outs]) + ',)'
            return 'return ' + self.w.pyname(p.result or p.name)
        if not p.outs:
            return 'return None'
        return 'return (' + ', '.join(self.w.pyname(a) for a in p.outs) + ',)'

    def stmt(self, s):
        h = s[0][1] if s[0][0] == 'name' else ''
        # a numeric statement label is not supported
        if s[0][0] == 'int':
            raise F90Unsupported('statement label: %r' % (s,))
        # assignment takes precedence over keywords (a variable may be called `if`...) when a top-level '=' follows a designator
        kind, k = self._find_assign(s)
        if kind and not (h in ('if', 'do', 'where', 'forall') and s[1] == ('op', '(') and match_paren(s, 1) < k) and not (h == 'do' and len(s) > 2 and s[2] == ('op', '=')):
            return self.assign(s[:k], s[k + 1:], kind)
        if h == 'if':
            m = match_paren(s, 1)
            cond, _ = self.gen(Parser(s[2:m]).expr())
            rest = s[m + 1:]
            if rest == [('name', 'then')]:
                self.emit('if %s:' % cond)
                self.ind += 1
                self.emit('pass')
                self.blocks.append('if')
            else:
                self.emit('if %s:' % cond)
                self.ind += 1
                self.stmt(rest)
                self.ind -= 1
            return
        if h in ('else', 'elseif'):
            self.ind -= 1
            if h == 'elseif' or (len(s) > 1 and s[1] == ('name', 'if')):
                i0 = 1 if h == 'elseif' else 2
                m = match_paren(s, i0)
                cond, _ = self.gen(Parser(s[i0 + 1:m]).expr())
                self.emit('elif %s:' % cond)
            else:
                self.emit('else:')
            self.ind += 1
            self.emit('pass')
            return
        if h in ('end', 'endif', 'enddo', 'endselect', 'endassociate'):
            what = s[1][1] if (h == 'end' and len(s) > 1) else h[3:]
            b = self.blocks.pop()
            if what == 'select':
                if b[0] != 'select':
                    raise F90Unsupported('end select closes %r' % (b,))
                self.ind -= 1 if b[2] else 0
                if not b[2]:
                    pass
                return
            if what == 'associate':
                for nm, old in b[1]:
                    if old is None:
                        self.vars.pop(nm, None)
                    else:
                        self.vars[nm] = old
                return
            self.ind -= 1
            return
        if h == 'do':
            self.blocks.append('do')
            if len(s) == 1:
                self.emit('while True:')
            elif s[1] == ('name', 'while'):
                m = match_paren(s, 2)
                cond, _ = self.gen(Parser(s[3:m]).expr())
                self.emit('while %s:' % cond)
            else:
                var = s[1][1]
                parts = split_top(s[3:])
                cs = [self.gen(Parser(list(x)).expr())[0] for x in parts]
                vv, is_g = self.lookup(var)
                py = self.w.pyname(var)
                if is_g:
                    self.globals_written.add(py)
                if len(cs) == 2:
                    self.emit('for %s in range(%s, (%s) + 1):' % (py, cs[0], cs[1]))
                else:
                    self.emit('for %s in _do(%s, %s, %s):' % (py, cs[0], cs[1], cs[2]))
            self.ind += 1
            self.emit('pass')
            return
        if h == 'exit':
            return self.emit('break')
        if h == 'cycle':
            return self.emit('continue')
        if h == 'return':
            return self.emit(self.ret())
        if h == 'continue':
            return self.emit('pass')
        if h == 'stop':
            return self.emit('raise F90Stop(%r)' % ' '.join(str(t[1]) for t in s[1:]))
        if h == 'call':
            return self.call(s[1:])
        if h == 'select':
            m = match_paren(s, 2)
            if s[1] == ('name', 'type'):
                inner = s[3:m]
                if ('op', '=>') in inner:
                    raise F90Unsupported('select type with associate name')
                c, _ = self.gen(Parser(inner).expr())
                self.blocks.append(['select', ('type', c, inner), False])
            else:
                c, T = self.gen(Parser(s[3:m]).expr())
                t = self.newtmp()
                self.emit('%s = %s' % (t, ('_s(%s)' % c) if T[0] == 'char' else c))
                self.blocks.append(['select', ('case', t, None), False])
            return
        if h in ('case', 'type', 'class') and self.blocks and self.blocks[-1][0] == 'select':
            b = self.blocks[-1]
            kw = 'elif' if b[2] else 'if'
            if b[2]:
                self.ind -= 1
            if b[1][0] == 'type':
                if s[:2] == [('name', 'class'), ('name', 'default')]:
                    self.emit('else:' if b[2] else 'if True:')
                else:
                    m = match_paren(s, 2)
                    tn = s[3][1]
                    var = b[1][2][0][1]
                    self.emit('%s isinstance(%s, T_%s):' % (kw, b[1][1], tn))
                    # inside the guard the selector has the guarded type
                    if var in self.vars:
                        nv = Var(var, 'type', tn)
                        nv.dummy = self.vars[var].dummy
                        self.vars[var] = nv
            elif s[1:] == [('name', 'default')]:
                self.emit('else:' if b[2] else 'if True:')
            else:
                m = match_paren(s, 1)
                conds = []
                for item in split_top(s[2:m]):
                    if ('op', ':') in item:
                        lo, hi = split_top(item, ':')
                        cc = []
                        if lo:
                            cc.append('%s <= %s' % (self.gen(Parser(lo).expr())[0], b[1][1]))
                        if hi:
                            cc.append('%s <= %s' % (b[1][1], self.gen(Parser(hi).expr())[0]))
                        conds.append('(' + ' and '.join(cc) + ')')
                    else:
                        conds.append('%s == %s' % (b[1][1], self.gen(Parser(item).expr())[0]))
                self.emit('%s %s:' % (kw, ' or '.join(conds)))
            b[2] = True
            self.ind += 1
            self.emit('pass')
            return
        if h == 'associate':
            m = match_paren(s, 1)
            saved = []
            for item in split_top(s[2:m]):
                nm = item[0][1]
                c, T = self.gen(Parser(item[2:]).expr())
                py = self.w.pyname(nm)
                self.emit('%s = %s' % (py, c))
                saved.append((nm, self.vars.get(nm)))
                nv = Var(nm, T[0], T[1], T[2], dims=[(None, None)] * T[2])
                self.vars[nm] = nv
            self.blocks.append(('associate', saved))
            return
        if h == 'allocate':
            m = match_paren(s, 1)
            for item in split_top(s[2:m]):
                if item and item[0][0] == 'name' and len(item) > 1 and item[1] == ('op', '=') and item[0][1] in ('stat', 'source', 'mold', 'errmsg'):
                    raise F90Unsupported('allocate with %s=' % item[0][1])
                node = Parser(list(item)).expr()
                segs = list(node[1])
                last_name, last_args = segs[-1]
                base_segs = segs[:-1] + [(last_name, None)]
                tgt, T = self.gen_des(base_segs, target=True)
                if last_args is None:
                    if T[0] != 'type':
                        raise F90Unsupported('allocate of a scalar %s' % tgt)
                    self.emit('%s = T_%s()' % (tgt, T[1]))
                    continue
                ext = []
                for a in last_args:
                    if a[0] == 'slice':
                        lo, hi = self.gen(a[1])[0], self.gen(a[2])[0]
                        # a lower bound above 1: elements 1..lo-1 are allocated too and never referenced (subscripts stay 1-based)
                        ext.append('_ub(%s, %s)' % (lo, hi))
                    else:
                        ext.append(self.gen(a)[0])
                tcls = 'T_%s' % T[1] if T[0] == 'type' else 'None'
                self.emit('%s = _newarr((%s,), %r, %s)' % (tgt, ', '.join(ext), T[0], tcls))
            return
        if h in ('deallocate', 'nullify'):
            m = match_paren(s, 1)
            for item in split_top(s[2:m]):
                if len(item) > 1 and item[1] == ('op', '='):
                    continue
                tgt, _ = self.gen_des(Parser(list(item)).expr()[1], target=True)
                self.emit('%s = None' % tgt)
            return
        if h == 'write':
            m = match_paren(s, 1)
            items = [self.gen(Parser(list(x)).expr())[0] for x in split_top(s[m + 1:]) if x]
            return self.emit('_records.append((%s,))' % ', '.join(items)) if items else self.emit('pass')
        if h == 'print':
            return self.emit('pass')
        if h in ('use', 'implicit', 'save', 'intent', 'format'):
            return
        d = parse_decl(s)
        if d is not None:
            raise F90Unsupported('declaration after the first executable statement in %s: %r' % (self.p.name, s))
        raise F90Unsupported('statement %r in %s' % (s, self.p.name if self.p else '?'))

    @staticmethod
    def _find_assign(s):
        depth = 0
        if s[0][0] != 'name':
            return None, -1
        for k, tk in enumerate(s):
            if tk[0] == 'op' and tk[1] in ('(', '(/', '['):
                depth += 1
            elif tk[0] == 'op' and tk[1] in (')', '/)', ']'):
                depth -= 1
            elif depth == 0 and tk in (('op', '='), ('op', '=>')):
                return tk[1], k
            elif depth == 0 and tk[0] in ('name',) and k > 0 and s[k - 1] not in (('op', '%'),) :
                # two names in a row at depth 0 before any '=': a keyword statement
                return None, -1
        return None, -1

    def assign(self, lhs, rhs, kind):
        node = Parser(list(lhs)).expr()
        if node[0] != 'des':
            raise F90Unsupported('assignment target %r' % (lhs,))
        tgt, Tl = self.gen_des(node[1], target=True)
        rc, Tr = self.gen(Parser(list(rhs)).expr())
        if kind == '=>':
            return self.emit('%s = %s' % (tgt, rc))
        if Tl[0] == 'type' and Tr[0] == 'type' and Tl[2] == 0 and Tr[2] == 0 and Tl[1] != Tr[1]:
            # defined assignment (interface assignment(=)): the specific whose dummies have these two types
            for cand in self.w.generics.get('assignment', []):
                key = self.rproc(cand)
                if key is None:
                    continue
                pc = self.w.analyse(key)
                if len(pc.args) != 2:
                    continue
                a0, a1 = pc.vars[pc.args[0]], pc.vars[pc.args[1]]
                if a0.base == 'type' and a1.base == 'type' and a1.tname == Tr[1] and (a0.tname == Tl[1] or self.w.extends(Tl[1], a0.tname)):
                    return self.call_proc(cand, [node, Parser(list(rhs)).expr()])
        last_args = node[1][-1][1]
        if Tl[2] > 0:
            if tgt.endswith(']'):
                return self.emit('%s = %s' % (tgt, rc))
            return self.emit('%s = _assign(%s, %s)' % (tgt, tgt, rc))
        if Tl[0] == 'type':
            return self.emit('%s = _copy(%s)' % (tgt, rc))
        if Tl[0] == 'real' and Tr[0] in ('int', 'unknown'):
            rc = 'float(%s)' % rc
        elif Tl[0] == 'int' and Tr[0] in ('real', 'unknown'):
            rc = 'int(%s)' % rc
        elif Tl[0] == 'char':
            rc = '_s(%s)' % rc
        self.emit('%s = %s' % (tgt, rc))

    def call(self, toks):
        node = Parser(list(toks)).expr()
        segs = node[1]
        name, args = segs[-1]
        args = args or []
        if len(segs) > 1:
            # type-bound procedure or procedure-pointer component
            base_code, T = self.gen_des(segs[:-1])
            comp = self.w.comp(T[1], name) if T[0] == 'type' else None
            if comp is not None and comp.base == 'proc':
                t = self.newtmp()
                self.emit('%s = %s' % (t, base_code))
                cargs = ', '.join([t] + [self.gen(a)[0] for a in args])
                return self.emit('%s.%s(%s)' % (t, self.w.attr(name), cargs))
            bp = self.w.bound(T[1], name) if T[0] == 'type' else None
            if bp is None:
                raise F90Unsupported('call %r' % (toks,))
            return self.call_proc(bp, [('des', segs[:-1])] + list(args))
        v, _ = self.lookup(name)
        if v is not None:  # dummy procedure / procedure pointer variable
            cargs = ', '.join(self.gen(a)[0] for a in args)
            return self.emit('%s(%s)' % (self.w.pyname(name), cargs))
        target = name
        if name in self.w.generics and name not in self.w.byname:
            target = self.w.generics[name][0]
        if target in _SUB_INTRINSICS:
            return self.emit(_SUB_INTRINSICS[target](self, args))
        return self.call_proc(target, list(args))

    def call_proc(self, target, args):
        name = target
        target = self.rproc(target)
        if target is None:
            raise F90Unsupported('subroutine %s not found' % name)
        p = self.w.analyse(target)
        self.deps.add(target)
        codes = [self.gen(a)[0] for a in args]
        callc = '%s(%s)' % (target, ', '.join(codes))
        if not p.outs:
            return self.emit(callc)
        # positions of the returned dummies among the actual arguments
        tg = []
        for o in p.outs:
            pos = p.args.index(o)
            a = None
            if pos < len(args) and args[pos][0] != 'kw':
                a = args[pos]
            else:
                for x in args:
                    if x[0] == 'kw' and x[1] == o:
                        a = x[2]
            if a is not None and a[0] == 'des':
                try:
                    v0, _ = self.lookup(a[1][0][0])
                    if v0 is None:
                        raise F90Unsupported('x')
                    tg.append(self.gen_des(a[1], target=True)[0])
                    continue
                except F90Unsupported:
                    pass
            tg.append('_')
        if all(x == '_' for x in tg):
            return self.emit(callc)
        self.emit('(%s,) = %s' % (', '.join(tg), callc))


# ------------------------------------------------------------------------------------------------ intrinsics
def _args_codes(cg, args):
    return [cg.gen(a) for a in args]


def _simple(fmt, Tfun):
    def f(cg, args):
        ac = _args_codes(cg, args)
        return fmt.format(*[c for c, _ in ac], all=', '.join(c for c, _ in ac)), Tfun([T for _, T in ac])
    return f


def _same(Ts):
    return Ts[0]


def _promote(Ts):
    T = Ts[0]
    for x in Ts[1:]:
        T = _arith_T(T, x)
    return T


def _realT(Ts):
    return ('real', None, Ts[0][2])


def _intT(Ts):
    return ('int', None, Ts[0][2] if Ts else 0)


def _scalar(base):
    return lambda Ts: (base, None, 0)


def _elem(Ts):
    return (Ts[0][0], Ts[0][1], 0)


def _size(cg, args):
    ac = _args_codes(cg, args)
    if len(ac) == 1:
        return '_size(%s)' % ac[0][0], ('int', None, 0)
    return '%s.shape[(%s) - 1]' % (ac[0][0], ac[1][0]), ('int', None, 0)


def _real_conv(cg, args):
    c, T = cg.gen(args[0])
    return ('_tofloat(%s)' % c), ('real', None, T[2])


def _int_conv(cg, args):
    c, T = cg.gen(args[0])
    return ('_toint(%s)' % c), ('int', None, T[2])


def _tiny(cg, args):
    return '2.2250738585072014e-308', ('real', None, 0)


def _huge(cg, args):
    _, T = cg.gen(args[0])
    return ('2147483647', ('int', None, 0)) if T[0] == 'int' else ('1.7976931348623157e308', ('real', None, 0))


def _sum(cg, args):
    ac = _args_codes(cg, args)
    if len(ac) > 1:
        return '_sumdim(%s, %s)' % (ac[0][0], ac[1][0]), (ac[0][1][0], None, ac[0][1][2] - 1)
    return '_sum(%s)' % ac[0][0], (ac[0][1][0], None, 0)


def _matmul(cg, args):
    (a, Ta), (b, Tb) = _args_codes(cg, args)
    return '_matmul(%s, %s)' % (a, b), ('real', None, Ta[2] + Tb[2] - 2)


_INTRINSICS = {
    'sqrt': _simple('_sqrt({0})', _realT), 'abs': _simple('abs({0})', _same), 'max': _simple('max({all})', _promote), 'min': _simple('min({all})', _promote),
    'dot_product': _simple('_dot({0}, {1})', lambda Ts: (_arith_T(Ts[0], Ts[1])[0], None, 0)), 'sum': _sum, 'matmul': _matmul, 'size': _size,
    'mod': _simple('_mod({0}, {1})', _promote), 'modulo': _simple('({0} % {1})', _promote), 'sign': _simple('_sign({0}, {1})', _same),
    'trim': _simple('_s({0})', _scalar('char')), 'adjustl': _simple('{0}.lstrip()', _scalar('char')), 'len_trim': _simple('len(_s({0}))', _scalar('int')),
    'len': _simple('len({0})', _scalar('int')), 'index': _simple('({0}.find({1}) + 1)', _scalar('int')),
    'ishft': _simple('_ishft({0}, {1})', _intT), 'iand': _simple('(({0}) & ({1}))', _intT), 'ior': _simple('(({0}) | ({1}))', _intT),
    'ieor': _simple('(({0}) ^ ({1}))', _intT), 'real': _real_conv, 'dble': _real_conv, 'float': _real_conv, 'int': _int_conv,
    'nint': _simple('_nint({0})', _intT), 'floor': _simple('math.floor({0})', _intT), 'ceiling': _simple('math.ceil({0})', _intT),
    'tiny': _tiny, 'huge': _huge, 'epsilon': lambda cg, a: ('2.220446049250313e-16', ('real', None, 0)),
    'maxval': _simple('_maxval({0})', _elem), 'minval': _simple('_minval({0})', _elem),
    'associated': lambda cg, a: ('(%s is not None)' % cg.gen(a[0])[0], ('logical', None, 0)) if len(a) == 1 else ('(%s is %s)' % (cg.gen(a[0])[0], cg.gen(a[1][2] if a[1][0] == 'kw' else a[1])[0]), ('logical', None, 0)),
    'allocated': _simple('({0} is not None)', _scalar('logical')), 'present': _simple('({0} is not None)', _scalar('logical')),
    'exp': _simple('math.exp({0})', _realT), 'log': _simple('math.log({0})', _realT), 'log10': _simple('math.log10({0})', _realT),
    'sin': _simple('math.sin({0})', _realT), 'cos': _simple('math.cos({0})', _realT), 'tan': _simple('math.tan({0})', _realT),
    'atan': _simple('math.atan({0})', _realT), 'atan2': _simple('math.atan2({0}, {1})', _realT), 'acos': _simple('math.acos({0})', _realT),
    'asin': _simple('math.asin({0})', _realT), 'merge': _simple('({0} if {2} else {1})', _same), 'null': lambda cg, a: ('None', ('unknown', None, 0)),
    'any': _simple('bool(np.any({0}))', _scalar('logical')), 'all': _simple('bool(np.all({0}))', _scalar('logical')),
    'count': _simple('int(np.count_nonzero({0}))', _scalar('int')), 'transpose': _simple('np.asfortranarray({0}.T)', _same),
}
_SUB_INTRINSICS = {
    'move_alloc': lambda cg, a: '%s = %s; %s = None' % (cg.gen_des(a[1][1], True)[0], cg.gen(a[0])[0], cg.gen_des(a[0][1], True)[0]),
}


# ------------------------------------------------------------------------------------------------ run-time helpers
def _newarr(shape, base, tcls=None):
    shape = tuple(int(x) for x in shape)
    if base == 'real':
        return np.zeros(shape, dtype=np.float64, order='F')
    if base == 'int':
        return np.zeros(shape, dtype=np.int64, order='F')
    if base == 'logical':
        return np.zeros(shape, dtype=bool, order='F')
    a = np.empty(shape, dtype=object, order='F')
    if base == 'type':
        for idx in np.ndindex(*shape):
            a[idx] = tcls()
    elif base == 'char':
        a[...] = ''
    return a


def _view1(a, n):
    """a rank-1 explicit-shape dummy: the first n elements of the actual argument (a view: writes reach the caller)"""
    if isinstance(a, np.ndarray) and a.ndim == 1 and a.shape[0] > n >= 0:
        return a[:n]
    return a


def _reshape(a, shape):
    if a is None:
        return None
    shape = tuple(int(x) for x in shape)
    if a.shape == shape:
        return a
    flat = a.reshape(-1, order='F') if a.ndim > 1 else a
    if -1 not in shape:
        n = int(np.prod(shape))
        if flat.size < n:  # the actual argument is shorter than the dummy's declared shape (legal while the rest is never referenced): keep the leading columns
            lead = int(np.prod(shape[:-1]))
            shape = shape[:-1] + (flat.size // lead,)
            n = lead * shape[-1]
        flat = flat[:n]
    r = flat.reshape(shape, order='F')
    if not np.shares_memory(r, a):
        raise F90Unsupported('sequence association would copy')
    return r


def _assign(dst, src):
    if dst is None:
        return np.array(src, copy=True, order='F')
    dst[...] = src
    return dst


def _copy(x):
    import copy
    return copy.copy(x)


def _idiv(a, b):
    a, b = int(a), int(b)
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def _div(a, b):
    if isinstance(a, (int, np.integer)) and isinstance(b, (int, np.integer)):
        return _idiv(a, b)
    return a / b


def _sq(x):
    return x * x


def _powi(x, n):
    n = int(n)
    if n < 0:
        return 1.0 / _powi(x, -n)
    r = 1 if isinstance(x, (int, np.integer)) else 1.0
    y = x
    while n:  # the multiplication chain of __builtin_powi (square and multiply)
        if n & 1:
            r = y if (isinstance(r, (int, float)) and r == 1 and not isinstance(x, (int, np.integer))) else r * y
        n >>= 1
        if n:
            y = y * y
    return r


def _powr(x, y):
    return math.pow(x, y) if not isinstance(x, np.ndarray) else np.power(x, y)


def _sqrt(x):
    return np.sqrt(x) if isinstance(x, np.ndarray) else math.sqrt(x)


def _dot(a, b):
    s = a[0] * b[0]
    for i in range(1, len(a)):
        s = s + a[i] * b[i]
    return s


def _sum(a):
    f = a.reshape(-1, order='F')
    s = f[0]
    for i in range(1, len(f)):
        s = s + f[i]
    return s


def _sumdim(a, dim):
    a = np.moveaxis(a, int(dim) - 1, 0)
    s = a[0].copy()
    for i in range(1, a.shape[0]):
        s = s + a[i]
    return s


def _matmul(a, b):
    if b.ndim == 1:
        out = np.zeros(a.shape[0])
        for i in range(a.shape[0]):
            s = 0.0
            for k in range(a.shape[1]):
                s = s + a[i, k] * b[k]
            out[i] = s
        return out
    out = np.zeros((a.shape[0], b.shape[1]), order='F')
    for j in range(b.shape[1]):
        for i in range(a.shape[0]):
            s = 0.0
            for k in range(a.shape[1]):
                s = s + a[i, k] * b[k, j]
            out[i, j] = s
    return out


def _ub(lo, hi):
    if int(lo) < 1:
        raise F90Unsupported('allocate with lower bound %d' % int(lo))
    return int(hi)


def _size(a):
    return int(a.size) if isinstance(a, np.ndarray) else len(a)


def _mod(a, b):
    if isinstance(a, (int, np.integer)) and isinstance(b, (int, np.integer)):
        return int(a) - int(b) * _idiv(a, b)
    return math.fmod(a, b)


def _sign(a, b):
    if isinstance(a, (int, np.integer)):
        return abs(int(a)) if b >= 0 else -abs(int(a))
    return math.copysign(abs(a), b)


def _s(x):
    return x.rstrip() if isinstance(x, str) else x


def _ishft(i, s):
    i, s = int(i), int(s)
    u = i & 0xFFFFFFFF
    u = (u << s) & 0xFFFFFFFF if s >= 0 else u >> (-s)
    return u - (1 << 32) if u & 0x80000000 else u


def _nint(x):
    return int(math.floor(x + 0.5)) if x >= 0 else -int(math.floor(-x + 0.5))


def _tofloat(x):
    return x.astype(np.float64) if isinstance(x, np.ndarray) else float(x)


def _toint(x):
    return x.astype(np.int64) if isinstance(x, np.ndarray) else int(x)


def _maxval(a):
    return a.max()


def _minval(a):
    return a.min()


def _do(a, b, c):
    a, b, c = int(a), int(b), int(c)
    return range(a, b + (1 if c > 0 else -1), c)


def _sl(lo, hi, st, lb):
    lb = int(lb)
    st = int(st)
    lo = None if lo is None else int(lo) - lb
    if hi is None:
        h = None
    else:
        h = int(hi) - lb + (1 if st > 0 else -1)
        if h < 0:
            h = None
    return slice(lo, h, st)


def _cat(parts):
    return np.concatenate([np.atleast_1d(p) for p in parts])


_RUNTIME = {k: v for k, v in list(globals().items()) if k.startswith('_') and callable(v) and not k.startswith('__')}
_RUNTIME.update({'np': np, 'math': math, 'F90Stop': F90Stop, 'F90Unsupported': F90Unsupported})
