import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _native_libraries():
    """The oracle and the host library are built on demand so a fresh checkout can run the CPU suite directly."""
    import subprocess
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    host = os.path.join(ROOT, "echorenderer_b200", "libecho_host.so")
    if not os.path.exists(host):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "echorenderer_b200", "csrc"), "../libecho_host.so"], check=True)


@pytest.fixture(scope="session")
def cornell():
    from echorenderer_b200 import host, scenes
    return host.prepare(scenes.cornell_box())


@pytest.fixture(scope="session")
def terrain_small():
    from echorenderer_b200 import host, scenes
    return host.prepare(scenes.terrain_scene(96, 48, 400))


@pytest.fixture(scope="session")
def mixed_small():
    from echorenderer_b200 import host, scenes
    return host.prepare(scenes.mixed_material_scene(rings=24, segments=24))


@pytest.fixture(scope="session")
def lights_small():
    from echorenderer_b200 import host, scenes
    return host.prepare(scenes.many_lights_scene(light_count=300, rings=16, segments=16))


@pytest.fixture(scope="session")
def coated_small():
    """The small mixed scene with CoatedDiffuse (Evaluation/Materials/CoatedDiffuse.cs) on the ground and on one blob."""
    import numpy as np
    from echorenderer_b200 import host, scenes
    description = scenes.mixed_material_scene(rings=24, segments=24)
    coated = np.concatenate([scenes.coated_diffuse((0.8, 0.5, 0.3), roughness=(0.3, 0.3), ior=1.5), scenes.coated_diffuse((0.9, 0.9, 0.9), roughness=(0.0, 0.6), ior=1.7)])
    first = len(description.materials)
    description.materials = np.concatenate([description.materials, coated])
    description.triangles["material"][description.triangles["material"] == 0] = first       # ground
    description.triangles["material"][description.triangles["material"] == 4] = first + 1   # the Lambertian blobs
    return host.prepare(description)


@pytest.fixture(scope="session")
def directional_small():
    """The small mixed scene lit by an AmbientLight, a DirectionalLight with a 2 degree cone and a delta DirectionalLight
    (Scenic/Lights/DirectionalLight.cs)."""
    import numpy as np
    from echorenderer_b200 import host, scenes
    description = scenes.mixed_material_scene(rings=24, segments=24)
    description.infinite_lights = np.concatenate([scenes.ambient_light((0.05, 0.05, 0.05)), scenes.directional_light((4.0, 3.6, 3.0), (50, 30, 0), angle=2.0, directly_visible=True),
                                                  scenes.directional_light((1.5, 1.5, 2.0), (70, -120, 0), angle=0.0)])
    return host.prepare(description)


@pytest.fixture(scope="session")
def textured_small():
    """Image textures in every material slot, normal mapping, sphere texture coordinates, an alpha cut-out (scenes.textured_scene)."""
    from echorenderer_b200 import host, scenes
    return host.prepare(scenes.textured_scene())


def sky_texture(height=32, width=64):
    """A latitude-longitude sky: a vertical gradient plus a small, very bright sun (what importance sampling is for)."""
    import numpy as np
    y, x = np.meshgrid((np.arange(height) + 0.5) / height, (np.arange(width) + 0.5) / width, indexing="ij")
    base = np.stack([0.3 + 0.5 * y, 0.4 + 0.5 * y, 0.6 + 0.4 * y], axis=-1)
    sun = np.exp(-((x - 0.3) ** 2 / 0.002 + (y - 0.8) ** 2 / 0.004))[..., None] * np.array([60.0, 55.0, 40.0])
    return np.concatenate([base + sun, np.ones((height, width, 1))], axis=-1)


@pytest.fixture(scope="session")
def environment_small():
    """The small mixed scene under an AmbientLight with a CylindricalTexture (importance-sampled environment map), rotated."""
    from echorenderer_b200 import host, scenes, structs
    description = scenes.mixed_material_scene(rings=24, segments=24)
    description.textures = [host.TextureDescription(sky_texture(), structs.FILTER_BILINEAR, structs.WRAPPER_REPEAT)]
    description.infinite_lights = scenes.environment_light(0, (1.0, 0.9, 0.8), (10, 40, 0))
    return host.prepare(description)


@pytest.fixture(scope="session")
def cubemap_small():
    """The small mixed scene under an AmbientLight with a Cubemap of six differently coloured gradient faces, rotated."""
    import numpy as np
    from echorenderer_b200 import host, scenes, structs
    description = scenes.mixed_material_scene(rings=24, segments=24)
    faces = []
    for k, colour in enumerate([(1.0, 0.3, 0.2), (0.2, 1.0, 0.3), (0.5, 0.7, 1.6), (0.4, 0.3, 0.2), (0.9, 0.9, 0.3), (0.8, 0.3, 0.9)]):
        y, x = np.meshgrid((np.arange(8) + 0.5) / 8, (np.arange(8) + 0.5) / 8, indexing="ij")
        shade = (0.4 + 0.6 * x * y + 0.2 * ((np.floor(x * 4) + np.floor(y * 4)) % 2))[..., None]
        faces.append(host.TextureDescription(np.concatenate([shade * np.array(colour), np.ones((8, 8, 1))], axis=-1), structs.FILTER_BILINEAR, structs.WRAPPER_CLAMP))
    description.textures = faces
    description.infinite_lights = scenes.cubemap_light(0, (0.8, 0.8, 0.8), (20, -30, 10))
    return host.prepare(description)
