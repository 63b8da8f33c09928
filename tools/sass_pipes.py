#!/usr/bin/env python
"""Classify the SASS instructions of a line range by issue pipe (developer tool).
usage: sass_pipes.py file.sass start end [start end ...]   (1-based inclusive line ranges)"""
import sys, re, collections
lines = open(sys.argv[1]).read().splitlines()
rng = list(map(int, sys.argv[2:]))
cnt = collections.Counter(); ops = collections.Counter()
for a, b in zip(rng[::2], rng[1::2]):
    for ln in lines[a - 1:b]:
        t = ln.split()
        if not t: continue
        op = t[1] if t[0].startswith('@') else t[0]
        op = op.rstrip(';')
        base = op.split('.')[0]
        if base in ('IMAD', 'FFMA', 'FMUL', 'FADD'): pipe = 'fma' + ('_wide' if '.WIDE' in op or '.HI' in op else '')
        elif base in ('LOP3', 'SHF', 'IADD3', 'ISETP', 'LEA', 'VIADD', 'SEL', 'MOV', 'PRMT', 'IABS', 'VIMNMX', 'IMNMX', 'PLOP3', 'P2R', 'R2P', 'FLO', 'POPC', 'CS2R', 'VOTE', 'VOTEU'): pipe = 'alu'
        elif base in ('LDG', 'STG', 'LDS', 'STS', 'LDC', 'ATOM', 'RED', 'REDG', 'ATOMG', 'SHFL', 'LDGSTS'): pipe = 'lsu'
        elif base.startswith('U') or base in ('LDCU', 'R2UR', 'S2UR', 'REDUX'): pipe = 'uniform'
        elif base in ('BRA', 'EXIT', 'BSSY', 'BSYNC', 'WARPSYNC', 'NOP', 'BAR', 'CALL', 'RET', 'S2R', 'YIELD'): pipe = 'ctrl'
        else: pipe = 'other:' + base
        cnt[pipe] += 1; ops[op] += 1
print(dict(cnt), 'total', sum(cnt.values()))
print(ops.most_common(30))
