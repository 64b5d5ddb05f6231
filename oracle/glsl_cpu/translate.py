"""Wrap one of the reference's Vulkan-GLSL compute shaders into a C++ translation unit for glsl_emu.hpp.

TEST INFRASTRUCTURE (part of oracle/).  The shader text is read from /root/reference at build time and is never
stored in this repository: the generated C++ goes to a temporary directory and only the compiled
oracle/_ref/libgsr_refshaders.so is kept (git-ignored).

What is rewritten -- declarations only; statements and expressions stay the shader author's:
  * `#[compute]`, `#version`, `#extension`, `#pragma unroll` lines are dropped;
  * `layout(local_size_*) in;` becomes the workgroup size of the dispatch function;
  * `layout(...) buffer|uniform Block { members };` becomes reference / array members of `struct Shader`, bound to the
    dispatch's buffers at their std430/std140 offsets; the push-constant block is copied from the dispatch's blob;
  * `layout(rgba32f...) uniform image2D x;` becomes an `image2D` member;
  * `shared T x[N];` / `shared T[N] x;` become members of the per-workgroup `Shader` object;
  * a global `const T X = e;` becomes `static constexpr`;
  * parameter qualifier `in` is dropped; a floating literal gets the `f` suffix (GLSL literals are 32-bit floats);
  * `T x = x;` (GLSL: the initialiser still sees the outer `x`, a block member) becomes `T x = this->x;`.
"""
from __future__ import annotations

import re

SCALAR_TYPES = {"uint", "int", "float", "vec2", "vec3", "vec4", "ivec2", "uvec2", "mat4"}


def strip_comments(src: str) -> str:
    src = re.sub(r"/\*.*?\*/", lambda m: "\n" * m.group(0).count("\n"), src, flags=re.S)
    return re.sub(r"//[^\n]*", "", src)


def suffix_float_literals(src: str) -> str:
    pat = re.compile(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][-+]?\d+)?|\d+[eE][-+]?\d+)(?![\w.])")
    return pat.sub(lambda m: m.group(1) + "f", src)


def parse_members(body: str):
    """[(type, name, array_expr | None | '')] -- '' means an unsized array."""
    out = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        m = re.fullmatch(r"(\w+)\s+(\w+)\s*(?:\[([^\]]*)\])?", decl)
        if not m:
            raise ValueError(f"cannot parse block member: {decl!r}")
        typ, name, arr = m.group(1), m.group(2), m.group(3)
        out.append((typ, name, None if arr is None else arr.strip()))
    return out


def translate(glsl: str, name: str) -> str:
    src = strip_comments(glsl)
    src = re.sub(r"^\s*#\[compute\]\s*$", "", src, flags=re.M)
    src = re.sub(r"^\s*#(version|extension|pragma)\b[^\n]*$", "", src, flags=re.M)
    src = suffix_float_literals(src)

    # workgroup size
    m = re.search(r"layout\s*\(([^)]*local_size_x[^)]*)\)\s*in\s*;", src)
    if not m:
        raise ValueError("no local_size layout")
    sizes = {"x": "1", "y": "1", "z": "1"}
    for part in m.group(1).split(","):
        k, v = part.split("=")
        sizes[k.strip()[-1]] = v.strip()
    src = src[: m.start()] + src[m.end():]

    ctor_init: list[str] = []   # reference members
    ctor_body: list[str] = []   # binds and copies
    n_bindings = 0

    def block_sub(m: re.Match) -> str:
        nonlocal n_bindings
        quals = m.group(1)
        members = parse_members(m.group(4))
        push = "push_constant" in quals
        bm = re.search(r"binding\s*=\s*(\d+)", quals)
        if not push and not bm:
            raise ValueError(f"block without binding: {quals}")
        binding = int(bm.group(1)) if bm else -1
        n_bindings = max(n_bindings, binding + 1)
        lines = []
        off = "size_t(0)"
        for typ, mname, arr in members:
            lay = f"::glsl::layout_of<{typ}>"
            off = f"::glsl::align_up({off}, {lay}::align)"
            off_name = f"_off_{mname}"
            lines.append(f"static constexpr size_t {off_name} = {off};")
            if push:
                if arr is not None:
                    raise ValueError("array in push-constant block")
                lines.append(f"{typ} {mname};")
                ctor_body.append(f"std::memcpy(&{mname}, static_cast<const char*>(_pc) + {off_name}, sizeof({typ}));")
                off = f"({off_name} + {lay}::size)"
            elif arr is None:
                if m.group(3) == "uniform":
                    lines.append(f"{typ} {mname};")
                    ctor_body.append(f"std::memcpy(&{mname}, static_cast<const char*>(_b[{binding}]) + {off_name}, sizeof({typ}));")
                else:
                    lines.append(f"{typ}& {mname};")
                    ctor_init.append(f"{mname}(*reinterpret_cast<{typ}*>(static_cast<char*>(_b[{binding}]) + {off_name}))")
                off = f"({off_name} + {lay}::size)"
            else:
                if typ not in SCALAR_TYPES and members[0][1] != mname:
                    raise ValueError("struct-typed array members are only supported at offset 0")
                lines.append(f"::glsl::buffer_array<{typ}> {mname};")
                if arr == "":
                    ctor_body.append(f"{mname}.bind(_b[{binding}], {off_name}, _s[{binding}]);")
                    off = f"({off_name})"
                else:
                    end = f"({off_name} + sizeof({typ}) * size_t({arr}))"
                    ctor_body.append(f"{mname}.bind(_b[{binding}], {off_name}, _s[{binding}] < {end} ? _s[{binding}] : {end});")
                    off = end
        return "\n".join(lines) + "\n"

    src = re.sub(r"layout\s*\(([^)]*)\)\s*((?:(?:restrict|readonly|writeonly|coherent)\s+)*)(buffer|uniform)\s+\w+\s*\{([^}]*)\}\s*;",
                 block_sub, src)

    def image_sub(m: re.Match) -> str:
        nonlocal n_bindings
        binding = int(re.search(r"binding\s*=\s*(\d+)", m.group(1)).group(1))
        n_bindings = max(n_bindings, binding + 1)
        nm = m.group(2)
        ctor_body.append(f"{nm}.texels = static_cast<float*>(_b[{binding}]); {nm}.width = int(_img_w); {nm}.height = int(_img_h);")
        return f"::glsl::image2D {nm};\n"

    src = re.sub(r"layout\s*\(([^)]*)\)\s*uniform\s+(?:(?:restrict|readonly|writeonly)\s+)*image2D\s+(\w+)\s*;", image_sub, src)

    # shared variables -> members of the per-workgroup object
    src = re.sub(r"\bshared\s+(\w+)\s*\[([^\]]*)\]\s*(\w+)\s*;", r"\1 \3[\2];", src)
    src = re.sub(r"\bshared\s+(\w+)\s+(\w+)\s*(\[[^\]]*\])?\s*;", lambda m: f"{m.group(1)} {m.group(2)}{m.group(3) or ''};", src)
    # global constants
    src = re.sub(r"^const\s+(\w+)\s+(\w+)\s*=", r"static constexpr \1 \2 =", src, flags=re.M)
    # parameter qualifier
    src = re.sub(r"([(,]\s*)in\s+(\w+\s+\w+\s*\[)", r"\1const \2", src)   # `in` array: a read-only copy
    src = re.sub(r"([(,]\s*)in\s+(?=\w)", r"\1", src)
    # `T x = x;` sees the outer x in GLSL
    src = re.sub(r"\b(\w+)\s+(\w+)\s*=\s*\2\s*;", r"\1 \2 = this->\2;", src)

    if re.search(r"\blayout\s*\(", src):
        raise ValueError("unhandled layout declaration left in shader " + name)

    init = (" : " + ", ".join(ctor_init)) if ctor_init else ""
    return f"""// generated by oracle/glsl_cpu/translate.py from the reference's {name}.glsl -- do not commit
#include <new>
#include "glsl_emu.hpp"
namespace glsl {{ namespace shader_{name} {{
struct Shader {{
{src}
    Shader(void* const* _b, const size_t* _s, const void* _pc, size_t _img_w, size_t _img_h){init} {{
        (void)_b; (void)_s; (void)_pc; (void)_img_w; (void)_img_h;
        {' '.join(ctor_body)}
    }}
    static void invoke(void* self) {{ static_cast<Shader*>(self)->main(); }}
}};
}} }}
extern "C" int refshader_{name}_bindings(void) {{ return {n_bindings}; }}
extern "C" void refshader_{name}_set_shared_fill(unsigned word) {{ glsl::shared_fill_word = word; }}
// one vkCmdDispatch(gx, gy, gz): workgroups in ascending order, a fresh set of `shared` variables per workgroup
extern "C" int refshader_{name}_dispatch(unsigned gx, unsigned gy, unsigned gz, void* const* buffers, const size_t* sizes,
                                         const void* push_constants, size_t image_width, size_t image_height) {{
    using namespace glsl;
    const uvec3 local(uint({sizes['x']}), uint({sizes['y']}), uint({sizes['z']}));
    for (unsigned z = 0; z < gz; ++z)
        for (unsigned y = 0; y < gy; ++y)
            for (unsigned x = 0; x < gx; ++x) {{
                // `shared` storage is uninitialised in GLSL: every word starts as glsl::shared_fill_word
                void* mem = ::operator new(sizeof(shader_{name}::Shader));
                fill_words(mem, sizeof(shader_{name}::Shader));
                shader_{name}::Shader* sh = new (mem) shader_{name}::Shader(buffers, sizes, push_constants, image_width, image_height);
                run_workgroup(uvec3(x, y, z), local, &shader_{name}::Shader::invoke, sh);
                sh->~Shader();
                ::operator delete(mem);
            }}
    return 0;
}}
"""


if __name__ == "__main__":
    import sys
    print(translate(open(sys.argv[1]).read(), sys.argv[2]))
