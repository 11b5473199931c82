#!/usr/bin/env python3
"""Mechanical GLSL-450 -> C++20 syntax rewrite of the reference compute shaders.

TEST INFRASTRUCTURE ONLY (part of the parity oracle, never of the product).

Reads the reference's shader sources *where they lie* (default
/root/reference/resources/shaders/source) and writes C++ text that can be
#include'd into the body of a struct (one struct instance == one shader
invocation) into oracle/_ref/.  No reference source is stored in this repo; the
output directory is git-ignored.

The rewrite is purely syntactic -- no arithmetic expression is changed:

  R1  `#version`, `layout(local_size...) in;`          -> dropped
  R2  `#include "x.glsl"`                              -> inlined (same rewrite applied)
  R3  float literals `1.0`, `.5`, `1e-4`               -> `1.0f` (GLSL literals are fp32,
      C++ ones are double; without this C++ would evaluate sub-expressions in fp64)
  R4  `inout T name` / `out T name` -> `T& name`;  leading `in ` qualifier dropped
  R5  multi-component swizzles `.zxy` -> `.zxy()` (glm function swizzles)
  R6  `layout(binding=N) uniform Block {..} name;`     -> `struct Block {..} name = bindUbo<Block>(N);`
  R7  `layout(binding=N, rgba8) uniform image2D name;` -> `image2D name = bindImage(N);`
  R8  `layout(std430, binding=N) readonly buffer B { T[] name; };`
                                                       -> `Ssbo<T> name = bindSsbo<T>(N);`
  R9  in definitions.glsl, `vec3 x;`/`vec4 x;` struct members get `alignas(16)` so the C++
      struct has the std430 layout the host uploads (checked by static_asserts in the driver)
  R10 the call `imageSize(` -> `imageSize_(` (main() declares a local of the same name, which
      in C++ would shadow the function inside its own initialiser)
  R12 a `vecN(...)` constructor call whose arguments themselves call random() is brace-initialised
      (`vecN{...}`): GLSL evaluates call arguments left to right (GLSL 4.50 spec, 6.1.1), C++ leaves the
      order of `f(a(), b(), c())` unspecified (GCC goes right to left) but guarantees it for braces
  R13 `--brute-force`: the active `if (hit_bvh(current_ray, rec))` of ray_color becomes the shader's own commented-out
      alternative one line above it, `if (hit_scene(current_ray, rec))` (ray-trace-compute.comp:322-323)
  R14 fragment shader interface: `layout(location=N) in T name;` / `out T name;` -> plain members; `layout(binding=N) uniform
      sampler2D name;` -> `sampler2D name = bindSampler(N);`; `#extension` lines dropped
  R15 `--enable-denoiser`: in main() of post-process-shader.frag the active `vec4 fragCol = texture(...)` is replaced by the
      shader's own commented-out line above it, `vec4 fragCol = t * smartDeNoise(...) + (1-t)*texture(...)` (:64-65)
  R11 optional overrides of `#define NUM_BOUNCES n` / `#define MAX_STACK_DEPTH n`
      (BASELINE configs need depth 4/8 and >16 stack for 1M-triangle trees; the verbatim
      variant keeps the shader's values)
"""
import argparse
import os
import re
import sys

FLOAT_LIT = re.compile(r'(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][-+]?\d+)?|\d+[eE][-+]?\d+)(?![\w.])')
SWIZZLE = re.compile(r'\.([xyzw]{2,4}|[rgba]{2,4})\b(?!\s*\()')


def strip_comments(src):
    src = re.sub(r'/\*.*?\*/', lambda m: '\n' * m.group(0).count('\n'), src, flags=re.S)
    src = re.sub(r'//[^\n]*', '', src)
    return src


def brace_random_ctor_args(src):
    out, i = [], 0
    for m in re.finditer(r'\bvec[234]\s*\(', src):
        if m.start() < i:
            continue
        depth, j = 1, m.end()
        while depth:
            depth += {'(': 1, ')': -1}.get(src[j], 0)
            j += 1
        inner = src[m.end():j - 1]
        if len(re.findall(r'\brandom\s*\(', inner)) >= 2:
            out.append(src[i:m.end() - 1] + '{' + inner + '}')
            i = j
    out.append(src[i:])
    return ''.join(out)


def rewrite(src, src_dir, is_definitions=False, overrides=None, brute_force=False, enable_denoiser=False):
    if brute_force:      # R13 (before comments are stripped: the alternative lives in one)
        src, n = re.subn(r'//\s*(if \(hit_scene\(current_ray, rec\)\) \{)\s*\n\s*if \(hit_bvh\(current_ray, rec\)\) \{', r'\1', src)
        if n != 1:
            raise SystemExit('brute-force: expected exactly one commented hit_scene alternative, found %d' % n)
    if enable_denoiser:  # R15
        src, n = re.subn(r'//\s*(vec4 fragCol = t \* smartDeNoise\([^\n]*\n)\s*vec4 fragCol = texture\(texSampler, fragTexCoord\);', r'\1', src)
        if n != 1:
            raise SystemExit('enable-denoiser: expected exactly one commented smartDeNoise line, found %d' % n)
    src = strip_comments(src)
    # R2 includes
    def inline(m):
        name = m.group(1)
        with open(os.path.join(src_dir, name)) as f:
            return rewrite(f.read(), src_dir, is_definitions=name.endswith('definitions.glsl'))
    src = re.sub(r'^[ \t]*#include\s+"([^"]+)"[ \t]*$', inline, src, flags=re.M)
    # R1
    src = re.sub(r'^[ \t]*#version[^\n]*$', '', src, flags=re.M)
    # R14
    src = re.sub(r'^[ \t]*#extension[^\n]*$', '', src, flags=re.M)
    src = re.sub(r'layout\s*\(\s*location\s*=\s*\d+\s*\)\s*(?:in|out)\s+(\w+)\s+(\w+)\s*;', r'\1 \2;', src)
    src = re.sub(r'layout\s*\(\s*binding\s*=\s*(\d+)\s*\)\s*uniform\s+sampler2D\s+(\w+)\s*;', r'sampler2D \2 = bindSampler(\1);', src)
    src = re.sub(r'layout\s*\(\s*local_size_x[^)]*\)\s*in\s*;', '', src)
    # R6 UBO block
    src = re.sub(r'layout\s*\(\s*binding\s*=\s*(\d+)\s*\)\s*uniform\s+(\w+)\s*\{(.*?)\}\s*(\w+)\s*;',
                 lambda m: 'struct %s {%s} %s = bindUbo<%s>(%s);' % (m.group(2), m.group(3), m.group(4), m.group(2), m.group(1)),
                 src, flags=re.S)
    # R7 storage images
    src = re.sub(r'layout\s*\(\s*binding\s*=\s*(\d+)\s*,\s*rgba8\s*\)\s*uniform\s+image2D\s+(\w+)\s*;',
                 r'image2D \2 = bindImage(\1);', src)
    # R8 SSBOs
    src = re.sub(r'layout\s*\(\s*std430\s*,\s*binding\s*=\s*(\d+)\s*\)\s*readonly\s+buffer\s+\w+\s*\{\s*(\w+)\s*\[\s*\]\s*(\w+)\s*;\s*\}\s*;',
                 r'Ssbo<\2> \3 = bindSsbo<\2>(\1);', src, flags=re.S)
    # R9 std430 member alignment
    if is_definitions:
        src = re.sub(r'^(\s*)(vec[34]\s+\w+\s*;)', r'\1alignas(16) \2', src, flags=re.M)
    # R4 parameter qualifiers
    src = re.sub(r'\b(?:inout|out)\s+(\w+)\s+(\w+)', r'\1& \2', src)
    src = re.sub(r'([(,]\s*)in\s+(\w+\s+\w+)', r'\1\2', src)
    # R10
    src = re.sub(r'\bimageSize\s*\(', 'imageSize_(', src)
    # R5 swizzles
    src = SWIZZLE.sub(r'.\1()', src)
    # R12
    src = brace_random_ctor_args(src)
    # R3 float literals (skip ones that already carry an f suffix: the look-ahead excludes \w)
    src = FLOAT_LIT.sub(r'\1f', src)
    # R11
    for name, val in (overrides or {}).items():
        src, n = re.subn(r'(#define\s+%s\s+)\d+' % name, r'\g<1>%d' % val, src)
        if n != 1:
            raise SystemExit('override %s: expected exactly one #define, found %d' % (name, n))
    return src


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--src-dir', default='/root/reference/resources/shaders/source')
    ap.add_argument('--shader', required=True, help='e.g. ray-trace-compute.comp')
    ap.add_argument('--out', required=True)
    ap.add_argument('--num-bounces', type=int)
    ap.add_argument('--max-stack-depth', type=int)
    ap.add_argument('--brute-force', action='store_true')
    ap.add_argument('--enable-denoiser', action='store_true')
    a = ap.parse_args()
    ov = {}
    if a.num_bounces is not None:
        ov['NUM_BOUNCES'] = a.num_bounces
    if a.max_stack_depth is not None:
        ov['MAX_STACK_DEPTH'] = a.max_stack_depth
    with open(os.path.join(a.src_dir, a.shader)) as f:
        out = rewrite(f.read(), a.src_dir, overrides=ov, brute_force=a.brute_force, enable_denoiser=a.enable_denoiser)
    with open(a.out, 'w') as f:
        f.write('// GENERATED by oracle/ref/glsl2cpp.py from %s -- do not commit\n' % a.shader)
        f.write(out)


if __name__ == '__main__':
    sys.exit(main())
