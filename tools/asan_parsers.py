#!/usr/bin/env python
"""tools/asan_parsers.py -- the host C layer (icw_plugin.c, icw_config.c) built with -fsanitize=address,undefined and fed the damaged headers
and configuration files of tests/test_plugin.py / tests/test_config_file.py (no GPU needed: parsing only).  Run as
    gcc -std=c99 -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -fPIC -shared -I include -o /tmp/asan/libicw_plugin.so \\
        in_cwave_b200/host/icw_plugin.c in_cwave_b200/host/icw_config.c -L in_cwave_b200 -licw_b200 -Wl,-rpath,$PWD/in_cwave_b200
    LD_PRELOAD=$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so) ASAN_OPTIONS=detect_leaks=0 python tools/asan_parsers.py
Any report of the sanitizers goes to stderr; round 2: 3600 headers and 5000 configuration files, none."""
import sys, tempfile, pathlib
ROOT = str(pathlib.Path(__file__).resolve().parent.parent); sys.path[:0] = [ROOT, ROOT + '/tests']
import numpy as np
from in_cwave_b200 import plugin
plugin.PLUGIN_PATH = pathlib.Path('/tmp/asan/libicw_plugin.so')
from oracle import pyoracle as po
from in_cwave_b200 import spec as S
import test_plugin as T
import test_config_file as TC
from util import random_spec
tmp=pathlib.Path(tempfile.mkdtemp())
n=0
for seed in range(300, 360):
    rng=np.random.default_rng(600+seed)
    for trial in range(60):
        is_cw=bool(rng.integers(0,2))
        if is_cw:
            d=dict(fmt=str(rng.choice(["cw_f64","cw_i16","cw_i16f32","cw_f32"])),n_channels=int(rng.integers(1,3)),sample_rate=96000)
            raw=rng.integers(0,256,size=int(rng.integers(2,400))*S.frame_bytes(d),dtype=np.uint8)
            blob=po.cwave_bytes(d,raw,version=int(rng.integers(1,3)))
        else:
            d=dict(fmt=str(rng.choice(["wav_u8","wav_i16","wav_i24","wav_i32","wav_f32"])),n_channels=int(rng.integers(1,3)),sample_rate=44100)
            raw=rng.integers(0,256,size=int(rng.integers(2,400))*S.frame_bytes(d),dtype=np.uint8)
            blob=po.wav_bytes(d,raw,extensible=bool(rng.integers(0,2)))
        p=tmp/("d.cwave" if is_cw else "d.wav")
        p.write_bytes(T._damage_header(rng,blob,is_cw))
        plugin.probe(p); n+=1
print('headers probed under ASan/UBSan:', n)
m=0
for seed in range(700, 900):
    rng=np.random.default_rng(4000+seed)
    d=random_spec(rng)
    f=tmp/'ref.cfg'
    assert po.ref_save_config(d,f)
    base=f.read_text().split("\n")
    while base and base[-1]=="": base.pop()
    for trial in range(25):
        lines=base
        for _ in range(int(rng.integers(1,4))): lines=TC._mutate(rng,lines)
        g=tmp/'damaged.cfg'; g.write_text("\n".join(lines)+"\n")
        plugin.load_config(g); m+=1
    plugin.save_config(tmp/'ours.cfg', d, sec_align=3)
    plugin.load_config(tmp/'ours.cfg')
print('config files loaded under ASan/UBSan:', m)
