#!/usr/bin/env python3
"""ORACLE -- TEST INFRASTRUCTURE ONLY (never imported, linked or run by lumen_b200/).

glsl2cpp: a mechanical source-to-source pass that turns one UNMODIFIED GLSL shader stage of the reference
(/root/reference/src/shaders/...) plus everything it #includes into a C++ fragment that compiles INSIDE a struct
body (`struct Prog : glslref::Stage { #include "gen/<stage>.inc" };`, see programs.cpp). GLSL is close enough to
C++ that the pass only has to rewrite what the two languages spell differently; every arithmetic expression,
statement and control-flow construct of the shader is emitted token for token:

  * `#include "x.glsl"`  -> expanded in place, once per stage (as glslang's include directive with the files' own
                            guards); host-shared headers (`*.h`: commons.h, path_commons.h, bdpt_commons.h) are
                            included unmodified at namespace scope by programs.cpp, straight from the reference tree
  * `#version`, `#extension` -> dropped
  * float literals       -> suffixed `f` (GLSL literals are single precision; C++ would compute in double)
  * `in/out/inout T p`   -> `T p` / `T& p` / `T& p` (no call on these paths aliases an in with an out argument)
  * swizzles `e.xyz`     -> `(glslref::swz<0,1,2>(e))`;   `e.xyz = r;` -> `glslref::swz_set<0,1,2>(e, r);`
  * `vecN(a(), b())` whose arguments contain more than one call of a function with side effects (the RNG) ->
                            evaluated into temporaries LEFT TO RIGHT (GLSL's order; C++ leaves it unspecified)
  * `layout(...)` interface declarations -> members bound through the stage environment (glslref::Stage::env):
        buffer_reference blocks  -> wrapper structs constructible from a 64-bit address (glslref::BufArray: element count from the
                                    harness's registry, out-of-range elements read as zero instead of faulting)
        uniform / buffer blocks  -> references to the memory bound at (set, binding)
        push_constant block      -> reference to the bound push-constant bytes
        image2D / sampler2D[] / accelerationStructureEXT -> handles of the environment
        rayPayloadEXT (location = N) -> plain member + an entry in payload_at(N) (what traceRayEXT's last argument names)
        rayPayloadInEXT / hitAttributeEXT -> reference to the incoming payload / the hit attributes
  * globals with initialisers become default member initialisers (run per invocation, in declaration order, as in GLSL)

Nothing else is touched. The generated files are written under oracle/_ref/gen/ (git-ignored: they are derived from
reference sources, which must not be copied into this repository).

usage: glsl2cpp.py <shader root> <stage file relative to root> <out.inc>
"""
import os
import re
import sys

HOST_HEADER = re.compile(r".*\.h$")

TOKEN_RE = re.compile(
    r"""
    (?P<ws>[ \t]+)
  | (?P<nl>\r?\n)
  | (?P<lcomment>//[^\n]*)
  | (?P<bcomment>/\*.*?\*/)
  | (?P<str>"(?:[^"\\\n]|\\.)*")
  | (?P<num>
        0[xX][0-9a-fA-F]+[uU]?
      | (?:\d+\.\d*|\.\d+)(?:[eE][-+]?\d+)?(?:lf|LF|[fF])?
      | \d+[eE][-+]?\d+(?:lf|LF|[fF])?
      | \d+[uU]?
    )
  | (?P<id>[A-Za-z_]\w*)
  | (?P<op><<=|>>=|\+\+|--|->|<<|>>|<=|>=|==|!=|&&|\|\||\+=|-=|\*=|/=|%=|&=|\|=|\^=|\#\#|.)
    """,
    re.X | re.S,
)


class Tok:
    __slots__ = ("k", "s")

    def __init__(self, k, s):
        self.k, self.s = k, s

    def __repr__(self):
        return f"{self.k}:{self.s!r}"


def tokenize(text):
    out = []
    pos = 0
    while pos < len(text):
        m = TOKEN_RE.match(text, pos)
        if not m:
            raise SyntaxError(f"cannot tokenize at {text[pos:pos+40]!r}")
        out.append(Tok(m.lastgroup, m.group()))
        pos = m.end()
    return out


def render(toks):
    return "".join(t.s for t in toks)


# --------------------------------------------------------------------------------------------- include expansion
def expand(root, rel, seen, out_lines):
    path = os.path.normpath(os.path.join(root, rel))
    if path in seen:
        return
    seen.add(path)
    with open(path, "r", encoding="utf-8", errors="replace") as fh:
        lines = fh.read().replace("\r\n", "\n").split("\n")
    out_lines.append(f'#line 1 "{path}"')
    for i, line in enumerate(lines, 1):
        s = line.strip()
        m = re.match(r'#\s*include\s+"([^"]+)"', s)
        if m:
            inc = m.group(1)
            if HOST_HEADER.match(inc):
                out_lines.append(f"// [glsl2cpp] host-shared header {inc}: included unmodified at namespace scope")
                continue
            expand(root, os.path.join(os.path.dirname(rel), inc), seen, out_lines)
            out_lines.append(f'#line {i + 1} "{path}"')
            continue
        if re.match(r"#\s*(version|extension)\b", s):
            out_lines.append("")
            continue
        out_lines.append(line)


# --------------------------------------------------------------------------------------------- token helpers
def is_code(t):
    return t.k not in ("ws", "nl", "lcomment", "bcomment")


def next_code(toks, i):
    """index of the next code token at or after i (len(toks) if none)"""
    while i < len(toks) and not is_code(toks[i]):
        i += 1
    return i


def prev_code(toks, i):
    while i >= 0 and not is_code(toks[i]):
        i -= 1
    return i


def match_forward(toks, i, open_s, close_s):
    """toks[i] is open_s; returns the index of the matching close_s"""
    depth = 0
    while i < len(toks):
        if toks[i].k == "op":
            if toks[i].s == open_s:
                depth += 1
            elif toks[i].s == close_s:
                depth -= 1
                if depth == 0:
                    return i
        i += 1
    raise SyntaxError("unbalanced " + open_s)


def match_backward(toks, i, open_s, close_s):
    depth = 0
    while i >= 0:
        if toks[i].k == "op":
            if toks[i].s == close_s:
                depth += 1
            elif toks[i].s == open_s:
                depth -= 1
                if depth == 0:
                    return i
        i -= 1
    raise SyntaxError("unbalanced " + close_s)


def T(text):
    return Tok("raw", text)


# --------------------------------------------------------------------------------------------- passes
def pass_float_literals(toks):
    for t in toks:
        if t.k != "num":
            continue
        s = t.s
        if s[:2] in ("0x", "0X"):
            continue
        if re.search(r"(lf|LF)$", s):
            t.s = s[:-2]  # a GLSL double literal stays a C++ double
            continue
        if s[-1] in "fF":
            continue
        if "." in s or "e" in s or "E" in s:
            t.s = s + "f"


SIDE_EFFECT_FUNCS = {"rand", "rand2", "rand3", "rand4", "mlt_rand"}
CTOR_TYPES = {"vec2", "vec3", "vec4", "uvec2", "uvec3", "uvec4", "ivec2", "ivec3", "ivec4"}


def split_args(toks, lo, hi):
    """token index ranges [a, b) of the comma-separated arguments inside toks[lo:hi]"""
    args, depth, start = [], 0, lo
    for i in range(lo, hi):
        t = toks[i]
        if t.k == "op":
            if t.s in "([{":
                depth += 1
            elif t.s in ")]}":
                depth -= 1
            elif t.s == "," and depth == 0:
                args.append((start, i))
                start = i + 1
    args.append((start, hi))
    return args


def pass_sequence_ctor_args(toks):
    """vecN(f(s), g(s)) with >= 2 arguments calling side-effecting functions: GLSL evaluates left to right."""
    i = 0
    while i < len(toks):
        t = toks[i]
        if t.k == "id" and t.s in CTOR_TYPES:
            j = next_code(toks, i + 1)
            if j < len(toks) and toks[j].k == "op" and toks[j].s == "(":
                close = match_forward(toks, j, "(", ")")
                args = split_args(toks, j + 1, close)
                n_side = sum(1 for a, b in args if any(x.k == "id" and x.s in SIDE_EFFECT_FUNCS for x in toks[a:b]))
                if n_side >= 2:
                    # inner constructors first (they are inside the argument ranges)
                    inner = toks[j + 1:close]
                    pass_sequence_ctor_args(inner)
                    args = split_args(inner, 0, len(inner))
                    parts = ["([&]{ "]
                    for n, (a, b) in enumerate(args):
                        parts.append(f"auto glsl_arg{n} = (" + render(inner[a:b]).strip() + "); ")
                    parts.append("return " + t.s + "(" + ", ".join(f"glsl_arg{n}" for n in range(len(args))) + "); }())")
                    toks[i:close + 1] = [T("".join(parts))]
        i += 1


def pass_param_qualifiers(toks):
    """in / out / inout in parameter lists."""
    i = 0
    while i < len(toks):
        t = toks[i]
        if t.k == "id" and t.s in ("in", "out", "inout"):
            # a parameter qualifier is preceded by '(' or ',' (optionally `const`) and followed by a type name
            p = prev_code(toks, i - 1)
            if p >= 0 and toks[p].k == "id" and toks[p].s == "const":
                p = prev_code(toks, p - 1)
            n = next_code(toks, i + 1)
            if p >= 0 and toks[p].k == "op" and toks[p].s in "(," and n < len(toks) and toks[n].k == "id":
                ref = t.s != "in"
                # drop the qualifier and the white space after it
                del toks[i]
                while i < len(toks) and toks[i].k == "ws":
                    del toks[i]
                n = next_code(toks, i)
                if toks[n].k == "id" and toks[n].s == "const":
                    n = next_code(toks, n + 1)
                if ref:
                    toks.insert(n + 1, T("&"))
                continue
        i += 1


SWZ = re.compile(r"^(?:[xyzw]{2,4}|[rgba]{2,4})$")
COMP = {"x": 0, "y": 1, "z": 2, "w": 3, "r": 0, "g": 1, "b": 2, "a": 3}
ASSIGN_OPS = {"=", "+=", "-=", "*=", "/="}


def postfix_start(toks, i):
    """toks[i] is the last token of a postfix expression; returns the index of its first token."""
    while True:
        t = toks[i]
        if t.k == "op" and t.s == ")":
            i = match_backward(toks, i, "(", ")")
            p = prev_code(toks, i - 1)
            if p >= 0 and (toks[p].k == "id" or (toks[p].k == "op" and toks[p].s in ")]")):
                i = p
                continue
            return i
        if t.k == "op" and t.s == "]":
            i = match_backward(toks, i, "[", "]")
            i = prev_code(toks, i - 1)
            continue
        if t.k in ("id", "raw"):
            p = prev_code(toks, i - 1)
            if p >= 0 and toks[p].k == "op" and toks[p].s == ".":
                i = prev_code(toks, p - 1)
                continue
            return i
        raise SyntaxError(f"cannot find the start of a swizzled expression near {render(toks[max(0, i-6):i+3])!r}")


def pass_swizzles(toks):
    i = 0
    while i < len(toks):
        t = toks[i]
        if t.k == "id" and SWZ.match(t.s) and i > 0:
            p = prev_code(toks, i - 1)
            n = next_code(toks, i + 1)
            if p >= 0 and toks[p].k == "op" and toks[p].s == "." and not (n < len(toks) and toks[n].k == "op" and toks[n].s == "("):
                base_end = prev_code(toks, p - 1)
                if toks[base_end].k == "num":  # 1.xx is not a swizzle
                    i += 1
                    continue
                start = postfix_start(toks, base_end)
                idx = ",".join(str(COMP[c]) for c in t.s)
                base = render(toks[start:base_end + 1])
                if n < len(toks) and toks[n].k == "op" and toks[n].s in ASSIGN_OPS:
                    # statement `base.swz op= rhs;`
                    semi = n
                    depth = 0
                    while not (toks[semi].k == "op" and toks[semi].s == ";" and depth == 0):
                        if toks[semi].k == "op" and toks[semi].s in "([{":
                            depth += 1
                        elif toks[semi].k == "op" and toks[semi].s in ")]}":
                            depth -= 1
                        semi += 1
                    rhs_toks = toks[n + 1:semi]
                    pass_swizzles(rhs_toks)
                    rhs = render(rhs_toks).strip()
                    op = toks[n].s
                    if op != "=":
                        rhs = f"(glslref::swz<{idx}>({base})) {op[0]} ({rhs})"
                    toks[start:semi] = [T(f"(glslref::swz_set<{idx}>({base}, {rhs}))")]
                    i = start + 1
                    continue
                toks[start:i + 1] = [T(f"(glslref::swz<{idx}>({base}))")]
                i = start + 1
                continue
        i += 1


def layout_args(toks, lo, hi):
    """{'binding': '3', 'scalar': True, ...} from the tokens between the parentheses of layout(...)"""
    d = {}
    for a, b in split_args(toks, lo, hi):
        s = render(toks[a:b]).strip()
        if "=" in s:
            k, v = s.split("=", 1)
            d[k.strip()] = v.strip()
        elif s:
            d[s] = True
    return d


def pass_layout(toks, payload_locs):
    i = 0
    while i < len(toks):
        t = toks[i]
        if t.k == "id" and t.s == "hitAttributeEXT":
            semi = i
            while not (toks[semi].k == "op" and toks[semi].s == ";"):
                semi += 1
            words = [x.s for x in toks[i + 1:semi] if is_code(x)]
            toks[i:semi + 1] = [T(f"{words[0]} {words[1]} = glsl_hit_attribs;")]
            i += 1
            continue
        if not (t.k == "id" and t.s == "layout"):
            i += 1
            continue
        j = next_code(toks, i + 1)
        close = match_forward(toks, j, "(", ")")
        la = layout_args(toks, j + 1, close)
        # declaration runs to the ';' at depth 0
        k = close + 1
        depth = 0
        brace_open = brace_close = None
        while not (toks[k].k == "op" and toks[k].s == ";" and depth == 0):
            if toks[k].k == "op" and toks[k].s == "{":
                if depth == 0:
                    brace_open = k
                depth += 1
            elif toks[k].k == "op" and toks[k].s == "}":
                depth -= 1
                if depth == 0:
                    brace_close = k
            k += 1
        semi = k
        head = [x.s for x in toks[close + 1:(brace_open if brace_open is not None else semi)] if is_code(x)]
        tail = [x.s for x in toks[(brace_close + 1) if brace_close is not None else semi:semi] if is_code(x)]
        set_no, binding = la.get("set", "0"), la.get("binding", "0")
        text = None
        if brace_open is not None:
            body = render(toks[brace_open + 1:brace_close]).strip()
            members = [m.strip() for m in body.split(";") if m.strip()]
            block_name = head[-1]
            if "buffer_reference" in la:
                # layout(buffer_reference, ...) [readonly] buffer Name { T member[]; };
                fields, inits, defaults = [], [], []
                if len(members) != 1:
                    raise SyntaxError(f"buffer_reference block {block_name}: expected one member, got {members}")
                m = re.match(r"^(.*?)\s+(\w+)\s*(\[\s*\])?$", members[0], re.S)
                ty, name = m.group(1), m.group(2)
                const = "const " if "readonly" in head else ""
                text = (f"struct {block_name} {{ glslref::BufArray<{const}{ty}> {name}; {block_name}() {{}} "
                        f"{block_name}(uint64_t glsl_addr) : {name}(glsl_addr) {{}} }};")
            elif "push_constant" in la:
                if tail:
                    raise SyntaxError("named push_constant instances are not handled")
                parts = []
                for mem in members:
                    m = re.match(r"^(.*?)\s+(\w+)$", mem, re.S)
                    parts.append(f"const {m.group(1)}& {m.group(2)} = *reinterpret_cast<const {m.group(1)}*>(env->push_constants);")
                if len(parts) != 1:
                    raise SyntaxError("push_constant block with more than one member")
                text = " ".join(parts)
            else:
                if tail:
                    raise SyntaxError(f"named interface block instance {tail} not handled")
                if len(members) != 1:
                    raise SyntaxError(f"interface block {block_name}: expected one member, got {members}")
                m = re.match(r"^(.*?)\s+(\w+)\s*(\[\s*\])?$", members[0], re.S)
                ty, name, arr = m.group(1), m.group(2), m.group(3)
                const = "const " if ("readonly" in head or "uniform" in head) else ""
                if arr:
                    text = f"{const}{ty}* {name} = reinterpret_cast<{const}{ty}*>(env->buffer({set_no}, {binding}));"
                else:
                    text = f"{const}{ty}& {name} = *reinterpret_cast<{const}{ty}*>(env->buffer({set_no}, {binding}));"
        else:
            words = head
            if "rayPayloadEXT" in words:
                ty, name = words[-2], words[-1]
                payload_locs.append((la["location"], name))
                text = f"{ty} {name};"
            elif "rayPayloadInEXT" in words:
                ty, name = words[-2], words[-1]
                text = f"{ty}& {name} = *reinterpret_cast<{ty}*>(glsl_incoming_payload);"
            elif "image2D" in words:
                text = f"glslref::image2D {words[-1]} = env->image({set_no}, {binding});"
            elif "accelerationStructureEXT" in words:
                text = f"glslref::accelerationStructureEXT {words[-1]} = env->accel({set_no}, {binding});"
            elif "sampler2D" in words:
                # `uniform sampler2D name[]` : tokens are name, [, ]
                name = [w for w in words if w not in ("uniform", "sampler2D", "[", "]")][-1]
                text = f"const glslref::sampler2D* {name} = env->samplers({set_no}, {binding});"
            elif "in" in words or "out" in words:
                # local_size_x etc. of compute stages: `layout(local_size_x = 1024, ...) in;`
                text = "/* " + render(toks[i:semi + 1]).replace("*/", "* /") + " */"
            else:
                raise SyntaxError(f"layout declaration not handled: {render(toks[i:semi+1])!r}")
        toks[i:semi + 1] = [T(text)]
        i += 1


def translate(root, rel):
    lines = []
    expand(root, rel, set(), lines)
    out = []
    payload_locs = []
    # preprocessor lines are translated too (their bodies hold literals and swizzles), but kept on their own lines
    text = "\n".join(lines)
    toks = tokenize(text)
    pass_float_literals(toks)
    pass_layout(toks, payload_locs)
    pass_param_qualifiers(toks)
    pass_sequence_ctor_args(toks)
    pass_swizzles(toks)
    out.append(render(toks))
    cases = " ".join(f"case {loc}: return &{name};" for loc, name in payload_locs)
    out.append(f"\n#line 1 \"glsl2cpp-epilogue\"\nvoid* payload_at(int glsl_loc) {{ switch (glsl_loc) {{ {cases} default: break; }} return nullptr; }}\n")
    return "".join(out)


def main():
    root, rel, dst = sys.argv[1:4]
    text = translate(root, rel)
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    with open(dst, "w") as fh:
        fh.write(f"// GENERATED by oracle/glslref/glsl2cpp.py from {os.path.join(root, rel)} -- do not edit, do not commit\n")
        fh.write(text)


if __name__ == "__main__":
    main()
