import sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from helpers import compare, render_gpu, render_oracle
from rusterix_b200 import scenes
for cfg in (scenes.cube(800,600,200,logo_size=256), scenes.teapot(960,540,60,logo_size=256), scenes.map_config(1920,1080,40,logo_size=256), scenes.dense(1280,720,40,patches=8)):
    r=cfg.rasterizer()
    g=render_gpu(r,cfg.scene,cfg.assets,cfg.width,cfg.height,cfg.tile_size)
    o=render_oracle(r,cfg.scene,cfg.assets,cfg.width,cfg.height,cfg.tile_size)
    print(cfg.name, compare(g,o,cfg.name))
