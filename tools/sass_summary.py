#!/usr/bin/env python3
"""Per-kernel instruction counts from `cuobjdump -sass spear_b200/libspear_particles.so` (widths of the global accesses,
griddepcontrol instructions, FFMA against FMUL+FADD, shared-memory atomics): `python tools/sass_summary.py > table.md`."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ['k_stream_fused', 'k_integrate_pass', 'k_expand_sorted', 'k_sort_pass', 'k_shade_vertices', 'k_expand_from_peers', 'k_sort_small', 'k_dead_append']


def main():
    txt = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "spear_b200", "libspear_particles.so")], capture_output=True, text=True).stdout
    print("| kernel | LDG 256 | STG 256 | LDG 128 | STG 128 | other LDG | other STG | FFMA | FMUL+FADD | ACQBULK / PREEXIT | ATOMS |")
    print("|---|---|---|---|---|---|---|---|---|---|---|")
    for f in re.split(r'\n\s*Function : ', txt)[1:]:
        name = f.split('\n', 1)[0].strip()
        short = [w for w in WANT if w in name]
        if not short:
            continue
        label = short[0] + re.sub(r'.*' + short[0], '', name)[:24]
        c = collections.Counter()
        for i in re.findall(r'/\*[0-9a-f]{4}\*/\s+(.*?);', f):
            c[i.split()[0] if not i.startswith('@') else i.split()[1]] += 1
        cnt = lambda pred: sum(v for k, v in c.items() if pred(k))
        l256, s256 = cnt(lambda k: k.startswith('LDG') and '.256' in k), cnt(lambda k: k.startswith('STG') and '.256' in k)
        l128, s128 = cnt(lambda k: k.startswith('LDG') and '.128' in k), cnt(lambda k: k.startswith('STG') and '.128' in k)
        print("| `%s` | %d | %d | %d | %d | %d | %d | %d | %d | %d / %d | %d |" % (
            label, l256, s256, l128, s128, cnt(lambda k: k.startswith('LDG')) - l256 - l128, cnt(lambda k: k.startswith('STG')) - s256 - s128,
            cnt(lambda k: k.startswith('FFMA')), cnt(lambda k: k.startswith('FMUL') or k.startswith('FADD')),
            cnt(lambda k: 'ACQBULK' in k), cnt(lambda k: 'PREEXIT' in k), cnt(lambda k: k.startswith('ATOMS'))))


if __name__ == "__main__":
    sys.exit(main())
