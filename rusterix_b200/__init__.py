"""rusterix_b200: a B200-native (sm_100a CUDA) implementation of Rusterix's rasterization hot path,
`Rasterizer::setup(..).rasterize(scene, pixels, width, height, tile_size, assets)`, behind the C ABI
of include/rxcuda.h.  This package is the host-side mirror of the reference's API for that path."""
from .types import (Assets, BBox, Batch2D, Batch3D, Chunk, CompiledLight, CompiledLinedef, MapMini, CullMode, GridShader, Light, LightType, MatVecMode,
                    PixelSource, PrimitiveMode, RenderMode, RepeatMode, SampleMode, Scene, Texture, Tile,
                    VGrayGradientShader)
from .camera import D3FirstPCamera, D3IsoCamera, D3OrbitCamera
from .rasterizer import DeviceContext, FrameBatch, Rasterizer
from ._lib import RxcError

__all__ = [
    "Assets", "BBox", "Batch2D", "Batch3D", "Chunk", "CompiledLight", "CompiledLinedef", "MapMini", "CullMode", "GridShader", "Light", "LightType", "MatVecMode",
    "PixelSource", "PrimitiveMode", "RenderMode", "RepeatMode", "SampleMode", "Scene", "Texture", "Tile",
    "VGrayGradientShader", "D3FirstPCamera", "D3IsoCamera", "D3OrbitCamera", "DeviceContext", "FrameBatch", "Rasterizer", "RxcError",
]
