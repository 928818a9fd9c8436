"""Host-side logic that needs no GPU: C-ABI surface, scene loaders, container round trips, error behaviour."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import cuda_pt_b200 as cpt
import oracle
from conftest import GOLDEN, ROOT, SCENES, have_meshes, have_vader
from cpts import read_cpts

needs_meshes = pytest.mark.skipif(not have_meshes(), reason="scene meshes not present (copied from the reference by build())")


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "cpt_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(cpt_b200_[a-z_0-9]+)\s*\(", hdr)))
    assert len(declared) >= 30
    L = C.CDLL(os.path.join(ROOT, "cuda-pt_b200", "libcpt_b200.so"))
    missing = [n for n in declared if not hasattr(L, n)]
    assert not missing, missing
    bound = set(cpt.lib()._signatures)
    assert set(declared) == bound, (set(declared) ^ bound)
    assert b"sm_100a" in cpt.lib().cpt_b200_version()


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under the package or include/ may reference it."""
    bad = []
    for base in ("cuda-pt_b200", "include"):
        for dp, _dn, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"liboracle|cpt_oracle|import oracle|oracle/|tools\.oracle", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_renderer_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    sc = cpt.Scene(os.path.join(GOLDEN, "cbox_wfpt_256.cpts"))
    with pytest.raises(cpt.CptError, match="no CUDA device"):
        cpt.WavefrontRenderer(sc, 0, 0)


def test_cpts_roundtrip_and_reader(tmp_path):
    src = os.path.join(GOLDEN, "cbox_mats_256.cpts")
    sc = cpt.Scene(src)
    out = str(tmp_path / "copy.cpts")
    sc.save(out)
    a, b = read_cpts(src), read_cpts(out)
    assert set(a) == set(b)
    for k in a:
        assert a[k].tobytes() == b[k].tobytes(), k  # bytes: node words hold ints that read as NaN floats
    cfg = sc.config
    assert (cfg["width"], cfg["height"], cfg["spp"], cfg["max_depth"], cfg["renderer"]) == (256, 256, 64, 8, 1)
    assert a["cie_xyz"].shape == (471, 4) and a["d65_spd"].shape == (531,)


def test_scene_crosses_the_abi_in_memory(tmp_path):
    """cpt_b200_scene_load_memory: the .cpts container handed over as a memory image (what the reference-side adapter
    integration/b200_wavefront.cuh does with the Scene the reference's loader built) gives the scene the file gives; the image is
    only read during the call; malformed images are refused like malformed files."""
    src = os.path.join(GOLDEN, "texroom_256.cpts")
    image = open(src, "rb").read()
    a, b = cpt.Scene(src), cpt.Scene(image)
    assert a.config == b.config and np.array_equal(a.camera, b.camera)
    out = str(tmp_path / "from_memory.cpts")
    b.save(out)
    x, y = read_cpts(src), read_cpts(out)
    assert set(x) == set(y) and all(x[k].tobytes() == y[k].tobytes() for k in x)
    with pytest.raises(cpt.CptError, match="bad magic|empty"):
        cpt.Scene(b"NOTASCENE" * 4)
    with pytest.raises(cpt.CptError):
        cpt.Scene(image[: len(image) // 3])


def test_scene_errors(tmp_path):
    with pytest.raises(cpt.CptError, match="cannot open"):
        cpt.Scene(str(tmp_path / "missing.cpts"))
    bad = tmp_path / "bad.cpts"
    bad.write_bytes(b"NOTASCENE" * 10)
    with pytest.raises(cpt.CptError, match="bad magic"):
        cpt.Scene(str(bad))
    with pytest.raises(cpt.CptError, match="cannot open scene file"):
        cpt.Scene(str(tmp_path / "missing.xml"))
    x = tmp_path / "v.xml"
    x.write_text('<scene version="1.1"><sensor/></scene>')
    with pytest.raises(cpt.CptError, match="version"):
        cpt.Scene(str(x))
    x.write_text('<scene version="1.2"><sensor><film/></sensor><brdf type="lambertian" id="a"/>'
                 '<shape type="sphere"><ref type="material" id="a"/></shape></scene>')
    with pytest.raises(cpt.CptError, match="TRIANGLE_ONLY"):
        cpt.Scene(str(x))
    x.write_text('<scene version="1.2"><sensor><film/></sensor><brdf type="lambertian" id="a">'
                 '<ref type="texture" id="t"/></brdf></scene>')
    with pytest.raises(cpt.CptError, match="unknown texture"):
        cpt.Scene(str(x))
    x.write_text('<scene version="1.2"><sensor></scene>')
    with pytest.raises(cpt.CptError, match="xml"):
        cpt.Scene(str(x))
    # truncated container
    data = open(os.path.join(GOLDEN, "cbox_wfpt_256.cpts"), "rb").read()
    t = tmp_path / "trunc.cpts"
    t.write_bytes(data[: len(data) // 2])
    with pytest.raises(cpt.CptError):
        cpt.Scene(str(t))


def test_pt_front_end_in_plain_c_builds_and_fails_loudly_without_a_gpu(tmp_path):
    """apps/pt.c (the reference's `pt <xml>` over the C ABI) compiles with gcc alone; on a box without a GPU it must
    exit non-zero with the library's message instead of falling back to anything."""
    import subprocess
    import torch
    import __graft_entry__ as g
    g._build_pt_app()
    exe = os.path.join(ROOT, "cuda-pt_b200", "pt")
    assert os.path.exists(exe)
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 1 and "Usage" in res.stderr
    res = subprocess.run([exe, str(tmp_path / "nope.xml")], capture_output=True, text=True)
    assert res.returncode == 2 and "cannot open" in res.stderr
    if not torch.cuda.is_available():
        res = subprocess.run([exe, os.path.join(GOLDEN, "cbox_wfpt_256.cpts"), str(tmp_path / "o.png")], capture_output=True, text=True)
        assert res.returncode == 3 and "no CPU fallback" in res.stderr and not (tmp_path / "o.png").exists()


def test_png_decoder_and_texture_conversion_match_an_independent_decoder(tmp_path):
    """The loader's own PNG reader (grey / RGB / RGBA / palette / 16-bit) against PIL, plus the texel conversion rules
    of src/impl/textures.cu:53-131 (byte/255, normal maps *2-1, luma for grey requests, 16->8 bit = high byte,
    the `data2[i + 1]` read of the second roughness map, roughness_to_alpha for ior maps)."""
    from PIL import Image
    sc = cpt.Scene(os.path.join(SCENES, "texroom_wfpt_256.xml"))
    out = tmp_path / "texroom.cpts"
    sc.save(str(out))
    d = read_cpts(str(out))
    td = np.frombuffer(d["tex_desc"].tobytes(), dtype=np.dtype([("w", "<i4"), ("h", "<i4"), ("c", "<i4"), ("p", "<i4"), ("off", "<u8")]))
    data = d["tex_data"].ravel()
    tex_dir = os.path.join(ROOT, "scenes", "assets", "tex")

    def get(i):
        t = td[i]
        o, n = int(t["off"]), int(t["w"]) * int(t["h"]) * int(t["c"])
        return data[o:o + n].reshape(int(t["h"]), int(t["w"]), int(t["c"]))

    def rgba(name, normal=False):
        im = Image.open(os.path.join(tex_dir, name))
        a = np.asarray(im.convert("RGB") if im.mode == "P" else im)
        if a.shape[2] == 3:
            a = np.concatenate([a, np.full(a.shape[:2] + (1,), 255, np.uint8)], -1)
        a = a.astype(np.float32) / np.float32(255)
        return a * np.float32(2) + np.float32(-1) if normal else a

    def grey(name):
        a = np.asarray(Image.open(os.path.join(tex_dir, name)))
        if a.dtype != np.uint8:
            return (a.astype(np.uint32) >> 8).astype(np.uint8)
        if a.ndim == 3:
            a = a.astype(np.int32)
            return ((a[..., 0] * 77 + a[..., 1] * 150 + 29 * a[..., 2]) >> 8).astype(np.uint8)
        return a

    # texture order = BSDF order x (diffuse, specular, glossy, normal, roughness), then emitters
    expect = [rgba("albedo.png"), rgba("normal.png", True), rgba("albedo_warm.png"), rgba("specular.png"), rgba("glossy.png"),
              rgba("glossy.png"), rgba("normal.png", True)]
    for i, e in enumerate(expect):
        assert np.array_equal(get(i), e), i
    rough = get(7)
    assert rough.shape[2] == 2
    assert np.array_equal(rough[..., 0], grey("rough_x.png").astype(np.float32) / np.float32(255))
    b = grey("rough_y.png").ravel()
    shifted = np.concatenate([b[1:], b[-1:]]).reshape(rough.shape[:2]).astype(np.float32) / np.float32(255)
    assert np.array_equal(rough[..., 1], shifted)
    ior = get(12)
    v = grey("ior16.png").astype(np.float32) / np.float32(255) * np.float32(1.5) + np.float32(1)
    x = np.log(np.maximum(v, np.float32(1e-3))).astype(np.float32)
    alpha = (np.float32(1.62142) + np.float32(0.819955) * x + np.float32(0.1734) * x * x + np.float32(0.0171201) * x * x * x +
             np.float32(0.000640711) * x * x * x * x)
    assert np.abs(ior[..., 0] - alpha).max() < 1e-6
    assert np.array_equal(get(len(td) - 1), rgba("envmap.png")) and np.array_equal(get(len(td) - 2), rgba("emission.png"))
    bs = np.frombuffer(d["bsdfs"].tobytes(), dtype=np.int32).reshape(len(d["bsdfs"]), -1)
    assert bs[0, -10:-5].tolist() == [0, -1, -1, 1, -1]  # floor: diffuse + normal
    assert sc.config["envmap_id"] == 3


@needs_meshes
def test_xml_loader_media_match_the_reference_dump(tmp_path):
    """<phase>/<medium> parsing, the medium + cullable bits of the per-primitive object word and the volumetric limits,
    against the scene the reference's loader produced for the same XML."""
    sc = cpt.Scene(os.path.join(SCENES, "medium_cbox_vpt_256.xml"))
    out = tmp_path / "m.cpts"
    sc.save(str(out))
    a, b = read_cpts(str(out)), read_cpts(os.path.join(GOLDEN, "medium_cbox_256.cpts"))
    assert np.array_equal(a["media"], b["media"])
    for k in (3, 4, 5, 6, 8, 9, 10, 12, 13, 19, 20, 21):  # depth limits, renderer, counts, max_volume, num_media, cam_vol
        assert a["config"][k] == b["config"][k], k
    assert sorted(np.unique(a["obj_idxs"] >> 20).tolist()) == sorted(np.unique(b["obj_idxs"] >> 20).tolist()) == [0, 1025, 1026, 1027]
    # per object: same medium word
    med = lambda d: {int(w & 0xfffff): int(w >> 20) for w in d["obj_idxs"]}
    assert med(a) == med(b)
    x = tmp_path / "g.xml"
    x.write_text('<scene version="1.2"><sensor><film/></sensor><phase type="isotropic" id="p"/>'
                 '<medium type="grid" id="m"><ref type="phase" id="p"/></medium></scene>')
    with pytest.raises(cpt.CptError, match="no density grid"):
        cpt.Scene(str(x))


VOLUMES = os.path.join(os.path.dirname(SCENES), "volume")


@needs_meshes
@pytest.mark.parametrize("xml,gold", [("grid_cbox_vpt_256.xml", "grid_cbox_256.cpts"), ("grid_cbox_rayleigh_vpt_256.xml", "grid_cbox_rayleigh_256.cpts")])
def test_xml_loader_grid_media_match_the_reference_dump(tmp_path, xml, gold):
    """Our NanoVDB reader (nvdb_read.cpp: leaves -> bricks, float map, root extrema, world box) + `<medium type="grid">`
    parsing against what the reference's loader + NanoVDB's own accessor produced for the same files: byte-identical
    media records, grid records (incl. majorant, box, mean density), brick tables, brick pools and blackbody table."""
    sc = cpt.Scene(os.path.join(SCENES, xml))
    out = tmp_path / "g.cpts"
    sc.save(str(out))
    a, b = read_cpts(str(out)), read_cpts(os.path.join(GOLDEN, gold))
    for k in ("media", "grids", "grid_bricks", "grid_data", "blackbody"):
        assert a[k].shape == b[k].shape and a[k].tobytes() == b[k].tobytes(), k


def test_cpts_grid_sections_are_validated(tmp_path):
    """A brick number past the end of the pool, or a grid record pointing past the brick table, is refused at load."""
    import struct
    data = bytearray(open(os.path.join(GOLDEN, "grid_cbox_rayleigh_256.cpts"), "rb").read())
    def section(name):
        n, = struct.unpack_from("<I", data, 8)
        off = 16
        for _ in range(n):
            nm = bytes(data[off:off + 24]).split(b"\0")[0].decode()
            elem, _r, count = struct.unpack_from("<IIQ", data, off + 24)
            off += 40
            if nm == name:
                return off, elem * count
            off += elem * count + (16 - (elem * count) % 16) % 16
        raise KeyError(name)
    off, nbytes = section("grid_bricks")
    bricks = np.frombuffer(bytes(data[off:off + nbytes]), dtype=np.int32)
    first = int(np.nonzero(bricks >= 0)[0][0])
    bad = bytearray(data)
    struct.pack_into("<i", bad, off + 4 * first, 1 << 20)
    p = tmp_path / "bad.cpts"
    p.write_bytes(bytes(bad))
    with pytest.raises(cpt.CptError, match="grid brick outside"):
        cpt.Scene(str(p))
    goff, _ = section("grids")
    bad = bytearray(data)
    struct.pack_into("<Q", bad, goff + 56 + 88, 1 << 40)  # density channel's brick_offset
    p.write_bytes(bytes(bad))
    with pytest.raises(cpt.CptError, match="brick table outside"):
        cpt.Scene(str(p))


def test_cpts_validation_of_externally_produced_containers(tmp_path):
    """.cpts files may come from other producers: fields that would hang or crash the device side are refused at
    load - a cached node with a zero skip offset (infinite loop in the cached walk) or pointing past the node array, texture
    descriptors with negative sizes, BSDF / emitter texture indices, the environment-map emitter index, emitter object references."""
    import struct
    src = os.path.join(GOLDEN, "texroom_256.cpts")
    data = bytearray(open(src, "rb").read())

    def section(name):
        n, = struct.unpack_from("<I", data, 8)
        off = 16
        for _ in range(n):
            nm = bytes(data[off:off + 24]).split(b"\0")[0].decode()
            elem, _r, count = struct.unpack_from("<IIQ", data, off + 24)
            off += 40
            if nm == name:
                return off, elem, count
            off += elem * count + (16 - (elem * count) % 16) % 16
        raise KeyError(name)

    def refused(patch, match):
        bad = bytearray(data)
        patch(bad)
        p = tmp_path / "bad.cpts"
        p.write_bytes(bytes(bad))
        with pytest.raises(cpt.CptError, match=match):
            cpt.Scene(str(p))

    coff, _, ccount = section("cache_nodes")
    assert ccount > 0
    w, = struct.unpack_from("<I", data, coff + 12)
    refused(lambda b: struct.pack_into("<I", b, coff + 12, w & ~0xff), "zero skip")
    refused(lambda b: struct.pack_into("<I", b, coff + 12, (0xffffff << 8) | (w & 0xff)), "past the node array")
    toff, telem, tcount = section("tex_desc")
    assert tcount > 0
    refused(lambda b: struct.pack_into("<i", b, toff, -4), "invalid dimensions")        # width
    refused(lambda b: struct.pack_into("<i", b, toff + 8, 3), "invalid dimensions")      # channels
    boff, belem, _ = section("bsdfs")
    refused(lambda b: struct.pack_into("<i", b, boff + 88, 1 << 20), "unknown texture")  # tex[0]
    eoff, eelem, ecount = section("emitters")
    refused(lambda b: struct.pack_into("<i", b, eoff + eelem + 4, -7), "invalid object reference")
    refused(lambda b: struct.pack_into("<i", b, eoff + eelem + 12, 1 << 20), "unknown texture")
    cfg, _, _ = section("config")
    refused(lambda b: struct.pack_into("<i", b, cfg + 4 * 11, 99), "environment-map emitter index")
    cpt.Scene(src).close()  # the unmodified file still loads


def test_nvdb_reader_rejects_what_it_cannot_read(tmp_path):
    src = open(os.path.join(VOLUMES, "puff64_density.nvdb"), "rb").read()
    def scene_with(data, kind="density"):
        (tmp_path / "v.nvdb").write_bytes(data)
        x = tmp_path / "g.xml"
        x.write_text('<scene version="1.2"><sensor><film/></sensor><phase type="isotropic" id="p"/><medium type="grid" id="m">'
                     '<string name="density" value="v.nvdb"/><ref type="phase" id="p"/></medium></scene>')
        return str(x)
    with pytest.raises(cpt.CptError, match="truncated"):
        cpt.Scene(scene_with(src[:100000]))
    with pytest.raises(cpt.CptError, match="not a NanoVDB file"):
        cpt.Scene(scene_with(b"\0" * 8 + src[8:]))
    zipped = bytearray(src)
    zipped[16 + 168] = 1  # metadata codec = ZIP
    with pytest.raises(cpt.CptError, match="compressed"):
        cpt.Scene(scene_with(bytes(zipped)))
    with pytest.raises(cpt.CptError, match="expected a float grid"):
        cpt.Scene(scene_with(open(os.path.join(VOLUMES, "puff64_albedo.nvdb"), "rb").read()))
    old = bytearray(src)
    old[8:12] = (31 << 21).to_bytes(4, "little")
    with pytest.raises(cpt.CptError, match="format version 31"):
        cpt.Scene(scene_with(bytes(old)))


def test_image_errors(tmp_path):
    x = tmp_path / "t.xml"
    notpng = tmp_path / "a.jpg"
    notpng.write_bytes(b"\xff\xd8\xff\xe0 not really a jpeg")
    x.write_text('<scene version="1.2"><sensor><film/></sensor><brdf type="lambertian" id="a">'
                 '<ref type="texture" id="t"/></brdf><texture id="t"><string name="diffuse" value="a.jpg"/></texture></scene>')
    with pytest.raises(cpt.CptError, match="JPEG"):
        cpt.Scene(str(x))
    (tmp_path / "a.hdr").write_bytes(b"#?RADIANCE\n\n-Y 1 +X 1\n\x80\x80\x80\x80")
    x.write_text(x.read_text().replace("a.jpg", "a.hdr"))
    with pytest.raises(cpt.CptError, match="HDR.*unsupported format"):
        cpt.Scene(str(x))
    (tmp_path / "a.bmp").write_bytes(b"BM" + bytes(64))
    x.write_text(x.read_text().replace("a.hdr", "a.bmp"))
    with pytest.raises(cpt.CptError, match="PNG"):
        cpt.Scene(str(x))
    x.write_text('<scene version="1.2"><sensor><film/></sensor><brdf type="lambertian" id="a">'
                 '<ref type="texture" id="t"/></brdf><texture id="t"><string name="diffuse" value="none.png"/></texture></scene>')
    with pytest.raises(cpt.CptError, match="cannot open"):
        cpt.Scene(str(x))


@needs_meshes
def test_xml_loader_matches_reference_loader(tmp_path):
    """Our XML loader vs the flat scene the reference's loader produced for the same XML: identical camera,
    configuration, materials, emitters and object table; an equivalent (not identical) BVH."""
    ours = cpt.Scene(os.path.join(SCENES, "cbox_wfpt_256.xml"))
    out = str(tmp_path / "ours.cpts")
    ours.save(out)
    a, r = read_cpts(out), read_cpts(os.path.join(GOLDEN, "cbox_wfpt_256.cpts"))
    assert np.array_equal(a["camera"].view(np.uint32), r["camera"].view(np.uint32))  # bit-identical DeviceCamera
    for k in (0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12):
        assert a["config"][k] == r["config"][k], k
    assert np.array_equal(a["bsdfs"][:, :108], r["bsdfs"][:, :108])
    assert np.array_equal(a["emitters"], r["emitters"])
    # object table: same materials / emitters / primitive counts / inverse areas (prim_offset differs: no SBVH split)
    oa, orf = a["objects"].view(np.int32).reshape(-1, 4), r["objects"].view(np.int32).reshape(-1, 4)
    assert np.array_equal(oa[:, [0, 2, 3]], orf[:, [0, 2, 3]])
    assert a["config"][13] == 36 and r["config"][13] == 53  # the reference's SBVH duplicates 17 references
    # same geometry: every triangle of ours appears in the reference's (duplicated) set
    def tri_set(d):
        t = np.concatenate([d["v0"], d["v1"], d["v2"]], axis=1)
        return {tuple(np.round(row, 5)) for row in t}
    assert tri_set(a) <= tri_set(r)
    # emitter primitive map points at the light's two triangles
    light = int(np.nonzero((oa[:, 0] & 0xFFFF) == 1)[0][0])
    for p in a["emitter_prims"]:
        assert (a["obj_idxs"][p] & 0xFFFFF) == light


@needs_meshes
def test_xml_loader_bvh_is_a_correct_reference_format_tree(tmp_path):
    """The BVH our loader emits is traversed by the oracle's restatement of the reference traversal and must find
    the brute-force closest hit (every BSDF scene: no SBVH, so the cached-phase quirk cannot drop hits...)"""
    ours = cpt.Scene(os.path.join(SCENES, "cbox_mats_wfpt_256.xml")) if False else cpt.Scene(os.path.join(SCENES, "cbox_wfpt_256.xml"))
    out = str(tmp_path / "ours.cpts")
    ours.save(out)
    sc = oracle.OracleScene(out)
    rs = np.random.RandomState(3)
    o = np.array([2.78, 2.73, -8.0], np.float32)
    for _ in range(600):
        d = np.array([rs.uniform(-0.35, 0.35), rs.uniform(-0.35, 0.35), 1.0], np.float32)
        d /= np.linalg.norm(d)
        p0, h0 = sc.closest_hit(o, d)
        p1, h1 = sc.closest_hit(o, d, brute=True)
        assert (p0 < 0) == (p1 < 0)
        if p1 >= 0:
            assert h0[0] <= h1[0] * (1 + 1e-6) + 1e-6 and h0[0] >= h1[0] * (1 - 1e-6) - 1e-6
    d = read_cpts(out)
    nodes = d["nodes"].reshape(-1, 2, 4)
    cnt = nodes[:, 1, 3].view(np.int32)
    base = nodes[:, 0, 3].view(np.int32)
    leaves = cnt > 0
    assert cnt[leaves].sum() == d["config"][13]                      # every primitive in exactly one leaf
    assert cnt[leaves].max() <= d["config"][17]                      # max_node_num honoured
    assert sorted(base[leaves]) == list(np.cumsum(np.r_[0, cnt[leaves]][:-1]))  # contiguous DFS-ordered ranges
    assert -cnt[0] == len(cnt)                                       # root skip offset spans the whole tree
    assert len(d["cache_nodes"]) <= 255


@needs_meshes
def test_render_of_our_loader_scene_matches_reference_statistically(tmp_path):
    """Oracle on OUR loader's scene vs the reference frame: different BVH and emitter-primitive order give a
    different noise realisation of the same integrand."""
    ours = cpt.Scene(os.path.join(SCENES, "cbox_mats_wfpt_256.xml")) if False else cpt.Scene(os.path.join(SCENES, "cbox_wfpt_256.xml"))
    out = str(tmp_path / "ours.cpts")
    ours.save(out)
    s, _ = oracle.OracleScene(out).render(16)
    ref = np.fromfile(os.path.join(GOLDEN, "cbox_wfpt_256_wfpt_f0064.bin"), dtype=np.float32).reshape(256, 256, 4)
    m = s / 16
    # robust block statistics (median luminance per 32x32 block); the blocks showing the tall box and the caustic
    # it throws on the ceiling are excluded: the reference drops some of that box's triangles (test_oracle.py),
    # a scene built by our loader does not
    bm = lambda x: np.median(x[:, :, :3].mean(-1).reshape(8, 32, 8, 32).transpose(0, 2, 1, 3).reshape(8, 8, -1), axis=-1)
    a, b = bm(m), bm(ref)
    mask = np.ones((8, 8), bool)
    mask[1:5, 1:4] = False
    assert np.abs(a[mask] - b[mask]).max() < 0.05
    assert abs(a[mask].mean() - b[mask].mean()) < 0.01


@pytest.mark.skipif(not have_vader(), reason="vader mesh not present")
def test_xml_loader_vader_scale():
    sc = cpt.Scene(os.path.join(SCENES, "vader_wfpt_1080p.xml"))
    cfg = sc.config
    assert (cfg["width"], cfg["height"], cfg["spp"], cfg["max_depth"]) == (1920, 1080, 1024, 16)
    assert cfg["num_prims"] == 320328 and cfg["num_emitters"] == 2 and cfg["num_bsdfs"] == 3
    assert cfg["num_nodes"] < 2 * cfg["num_prims"] and cfg["num_cache"] == 15
    cam = sc.camera
    assert abs(cam[16] - 0.19) < 1e-7 and abs(cam[17] - 5.4) < 1e-6  # thin lens
    sc.set_film(256, 128)
    assert sc.config["width"] == 256 and sc.camera[13] == 128.0 and sc.camera[14] == 64.0


def test_pair_layout_and_leaf_refinement_invariants(tmp_path):
    """build_pair_layout (the re-laid-out traversal structure) with and without leaf refinement on every golden scene:
    each primitive in exactly one leaf record, refined records inside their reference leaf's range and box, the
    permutation only moves primitives inside reference leaves.  Plain C++ against scene_host.cpp, no GPU."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "pair_layout_check")
    subprocess.run(["g++", "-std=c++17", "-O1", "-I" + os.path.join(root, "include"), "-I" + os.path.join(root, "cuda-pt_b200", "csrc"),
                    os.path.join(root, "tests", "cpp", "pair_layout_check.cpp"), os.path.join(root, "cuda-pt_b200", "csrc", "scene_host.cpp"), "-o", exe],
                   check=True)
    refined_any = 0
    for scene in ("cbox_wfpt_256", "cbox_mats_256", "texroom_256", "medium_cbox_256", "grid_cbox_256"):
        for refine in (0, 2, 4):
            out = subprocess.run([exe, os.path.join(GOLDEN, scene + ".cpts"), str(refine)], capture_output=True, text=True)
            assert out.returncode == 0, (scene, refine, out.stdout, out.stderr)
            fields = dict(zip(out.stdout.split()[0::2], out.stdout.split()[1::2]))
            assert int(fields["errors"]) == 0
            if refine == 0:
                assert int(fields["refined_leaves"]) == 0
            refined_any += int(fields["refined_leaves"])
    assert refined_any > 0
    # trees built by our own loader (binned SAH, best of four bin counts, insertion-based optimisation): same invariants,
    # every child box inside its parent's, shallow enough for the traversal stack, no leaf inside the cached levels
    if have_meshes():
        for xml in ("cbox_mats_wfpt_256.xml", "medium_cbox_vpt_256.xml", "texroom_wfpt_256.xml"):
            sc = cpt.Scene(os.path.join(SCENES, xml))
            path = str(tmp_path / "own.cpts")
            sc.save(path)
            out = subprocess.run([exe, path, "2"], capture_output=True, text=True)
            assert out.returncode == 0, (xml, out.stdout, out.stderr)
            fields = dict(zip(out.stdout.split()[0::2], out.stdout.split()[1::2]))
            assert int(fields["errors"]) == 0 and int(fields["nesting_violations"]) == 0 and int(fields["depth"]) <= 46, (xml, fields)
            d = read_cpts(path)
            nodes = d["nodes"].reshape(-1, 2, 4)
            cnt = nodes[:, 1, 3].copy().view(np.int32)
            cached_idx = (d["cache_nodes"][:, 3] >> 8).astype(np.int64)
            assert not (cnt[cached_idx] > 0).any(), xml  # no leaf among the fp16-cached nodes


REF_DUMP = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "ref_dump")


@needs_meshes
@pytest.mark.skipif(not os.path.exists(REF_DUMP), reason="reference harness (NanoVDB tools) not built")
def test_nvdb_reader_matches_nanovdb_accessor_incl_tiles(tmp_path):
    """Our .nvdb reader against NanoVDB's own host accessor (`ref_dump vdbdump`, no GPU) on a grid whose constant regions
    the NanoVDB builder pruned into tile values (60 lower-level tiles), with negative index coordinates, and on the
    ordinary puff: identical float map, index box, brick table and brick pool."""
    import subprocess
    prefix = str(tmp_path / "t")
    subprocess.run([REF_DUMP, "mkvdb", prefix, "32"], check=True, capture_output=True)
    mesh = lambda n: os.path.relpath(os.path.join(os.path.dirname(SCENES), "meshes", "cbox", n), str(tmp_path))
    for kind, min_tiles in (("blocks", 1), ("density", 0)):
        out = subprocess.run([REF_DUMP, "vdbdump", f"{prefix}_{kind}.nvdb", f"{prefix}_{kind}.cpts"], check=True, capture_output=True, text=True)
        tiles = int(out.stdout.split("tiles in tree:")[1].split()[0])
        assert tiles >= min_tiles
        x = tmp_path / f"{kind}.xml"
        x.write_text(f'''<scene version="1.2"><renderer type="vpt"/><sensor><film/></sensor><phase type="isotropic" id="p"/>
<medium type="grid" id="m"><string name="density" value="t_{kind}.nvdb"/><ref type="phase" id="p"/></medium>
<brdf type="lambertian" id="w"/><emitter type="area" id="a"><rgb name="emission" value="1,1,1"/><string name="bind_type" value="triangle"/></emitter>
<shape type="obj"><string name="filename" value="{mesh("light.obj")}"/><ref type="material" id="w"/><ref type="emitter" id="a"/></shape>
<shape type="obj"><string name="filename" value="{mesh("floor.obj")}"/><ref type="material" id="w"/></shape></scene>''')
        sc = cpt.Scene(str(x))
        sc.save(str(tmp_path / "own.cpts"))
        a, b = read_cpts(str(tmp_path / "own.cpts")), read_cpts(f"{prefix}_{kind}.cpts")
        ga, gb = a["grids"][0], b["grids"][0]
        assert ga[56:160].tobytes() == gb[56:160].tobytes()          # density channel: map, index box, background
        assert ga[20:24].tobytes() == gb[20:24].tobytes() and ga[32:56].tobytes() == gb[32:56].tobytes()  # majorant, world box
        assert np.array_equal(a["grid_bricks"], b["grid_bricks"]) and a["grid_data"].tobytes() == b["grid_data"].tobytes()
    # Vec3f grid with tiles as the albedo channel of a medium whose density is the ordinary puff
    out = subprocess.run([REF_DUMP, "vdbdump", f"{prefix}_blocks3.nvdb", f"{prefix}_blocks3.cpts"], check=True, capture_output=True, text=True)
    assert int(out.stdout.split("tiles in tree:")[1].split()[0]) >= 1
    x = tmp_path / "v3.xml"
    x.write_text(open(str(tmp_path / "density.xml")).read().replace('<string name="density" value="t_density.nvdb"/>',
                 '<string name="density" value="t_density.nvdb"/><string name="albedo" value="t_blocks3.nvdb"/>'))
    sc = cpt.Scene(str(x))
    sc.save(str(tmp_path / "own3.cpts"))
    a, b, d = read_cpts(str(tmp_path / "own3.cpts")), read_cpts(f"{prefix}_blocks3.cpts"), read_cpts(f"{prefix}_density.cpts")
    ga, gb = a["grids"][0], b["grids"][0]
    alb_a, alb_b = ga[160:264].copy(), gb[160:264].copy()   # albedo channel record; offsets differ (it follows the density channel)
    nb_d, nd_d = d["grid_bricks"].size, d["grid_data"].size
    assert alb_a[:88].tobytes() == alb_b[:88].tobytes()
    assert np.frombuffer(alb_a[88:104].tobytes(), dtype=np.uint64).tolist() == [nb_d, nd_d]
    assert np.array_equal(a["grid_bricks"][nb_d:], b["grid_bricks"]) and a["grid_data"][nd_d:].tobytes() == b["grid_data"].tobytes()


@pytest.mark.skipif(not os.path.exists(REF_DUMP), reason="reference harness not built")
def test_image_decoders_match_the_reference_decoder_byte_for_byte(tmp_path):
    """Our PNG, JPEG and Radiance-HDR decoders against the decoder the reference links (`ref_dump imgdump` = its stb_image
    instance, CPU only): 4:4:4 / 4:2:2 / 4:2:0 / 4:1:1 chroma, greyscale, restart markers, optimised Huffman tables, low and
    high quality, odd and tiny sizes, progressive files (spectral selection + successive approximation, with restart
    markers), 1 / 2 / 3 / 4 requested channels (1 / 2 from a colour JPEG = its luma plane)."""
    import subprocess
    Image = pytest.importorskip("PIL.Image")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    csrc = os.path.join(root, "cuda-pt_b200", "csrc")
    exe = str(tmp_path / "image_check")
    subprocess.run(["g++", "-std=c++17", "-O1", "-I" + csrc, os.path.join(root, "tests", "cpp", "image_check.cpp"),
                    os.path.join(csrc, "image_io.cpp"), os.path.join(csrc, "jpeg_read.cpp"), os.path.join(csrc, "hdr_read.cpp"), "-lz", "-o", exe],
                   check=True)
    rng = np.random.default_rng(2)
    def img(w, h, amp=100):
        y, x = np.mgrid[0:h, 0:w]
        a = np.stack([128 + amp * np.sin(x / 3.0) * np.cos(y / 4.0), 128 + amp * np.cos(x / 2.0 + y / 5.0), 128 + amp * np.sin((x + y) / 3.0)], -1)
        return Image.fromarray(np.clip(a + rng.normal(0, 25, a.shape), 0, 255).astype(np.uint8))
    cases = [("a444.jpg", img(67, 45), dict(quality=92, subsampling=0)), ("a422.jpg", img(67, 45), dict(quality=75, subsampling=1)),
             ("a420.jpg", img(67, 45), dict(quality=85, subsampling=2)), ("a411.jpg", img(97, 61, 140), dict(quality=80, subsampling="4:1:1")),
             ("grey.jpg", img(128, 96).convert("L"), dict(quality=80)), ("rst.jpg", img(200, 130), dict(quality=70, subsampling=2, restart_marker_blocks=3)),
             ("opt.jpg", img(33, 17), dict(quality=30, subsampling=2, optimize=True)), ("q100.jpg", img(97, 61, 140), dict(quality=100, subsampling=0)),
             ("hard.jpg", img(97, 61, 140), dict(quality=15, subsampling=2)), ("t1.jpg", img(1, 1), dict(quality=70)), ("t2.jpg", img(2, 3), dict(quality=70)),
             ("t17.jpg", img(17, 33), dict(quality=70)), ("p1.jpg", img(64, 64), dict(quality=80, progressive=True)),
             ("p2.jpg", img(97, 61, 140), dict(quality=90, progressive=True, subsampling=0)),
             ("p3.jpg", img(200, 130), dict(quality=50, progressive=True, subsampling=2, restart_marker_blocks=5)),
             ("p4.jpg", img(37, 29).convert("L"), dict(quality=75, progressive=True)),
             ("p5.jpg", img(33, 17), dict(quality=20, progressive=True, subsampling=1, optimize=True)), ("rgb.png", img(31, 19), {}), ("grey.png", img(31, 19).convert("L"), {}),
             ("rgba.png", img(31, 19).convert("RGBA"), {})]
    for name, im, kw in cases:
        path = str(tmp_path / name)
        im.save(path, **kw)
        for req in (1, 2, 3, 4):
            dump = str(tmp_path / (name + f".{req}.bin"))
            subprocess.run([REF_DUMP, "imgdump", path, str(req), dump], check=True, capture_output=True)
            out = subprocess.run([exe, path, str(req), dump], capture_output=True, text=True)
            assert out.returncode == 0 and "mismatching bytes 0 of" in out.stdout, (name, req, out.stdout)
    # Radiance .hdr through the 8-bit entry point (RGBE -> float -> gamma 1/2.2 bytes), run-length encoded and flat scanlines
    def write_hdr(path, a, rle):
        h, w = a.shape[:2]
        m = a.max(-1)
        px = np.zeros((h, w, 4), np.uint8)
        nz = m > 1e-32
        fr, ex = np.frexp(m[nz])
        px[nz, :3] = (a[nz] * (fr * 256.0 / m[nz])[:, None]).astype(np.uint8)
        px[nz, 3] = (ex + 128).astype(np.uint8)
        with open(path, "wb") as f:
            f.write(b"#?RADIANCE\nGAMMA=1\nFORMAT=32-bit_rle_rgbe\n\n" + f"-Y {h} +X {w}\n".encode())
            if not rle:
                f.write(px.tobytes())
                return
            for j in range(h):
                f.write(bytes([2, 2, w >> 8, w & 255]))
                for k in range(4):
                    row, i = px[j, :, k], 0
                    while i < w:
                        r = 1
                        while i + r < w and r < 127 and row[i + r] == row[i]:
                            r += 1
                        if r >= 3:
                            f.write(bytes([128 + r, row[i]]))
                            i += r
                        else:
                            n = 1
                            while i + n < w and n < 128 and not (i + n + 2 < w and row[i + n] == row[i + n + 1] == row[i + n + 2]):
                                n += 1
                            f.write(bytes([n]) + row[i:i + n].tobytes())
                            i += n
    def hdr_img(w, h):
        y, x = np.mgrid[0:h, 0:w]
        a = np.stack([np.exp(np.sin(x / 5.0) * 3), np.exp(np.cos(y / 4.0) * 2) * 0.5, (x + y) % 7 / 7.0 * 4], -1) * rng.uniform(0.5, 1.5, (h, w, 1))
        a[h // 2:h // 2 + 3, w // 3:w // 3 + 9] = 0.0
        a[:2, :12] = 2.5
        return a
    for name, (w, h), rle in (("rle.hdr", (64, 20), True), ("flat.hdr", (40, 11), False), ("tiny.hdr", (5, 3), False)):
        path = str(tmp_path / name)
        write_hdr(path, hdr_img(w, h), rle)
        for req in (1, 2, 3, 4):
            dump = str(tmp_path / (name + f".{req}.bin"))
            subprocess.run([REF_DUMP, "imgdump", path, str(req), dump], check=True, capture_output=True)
            out = subprocess.run([exe, path, str(req), dump], capture_output=True, text=True)
            assert out.returncode == 0 and "mismatching bytes 0 of" in out.stdout, (name, req, out.stdout)
    # hand-written PNGs: Adam7 interlacing and the sub-byte / 16-bit / palette + tRNS layouts PIL does not write
    import struct, zlib
    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xffffffff)
    def write_png(path, rows_fn, w, h, depth, ctype, interlace, palette=None, trns=None):
        raw = b""
        if interlace:
            xo, yo, xs, ys = [0, 4, 0, 2, 0, 1, 0], [0, 0, 4, 0, 2, 0, 1], [8, 8, 4, 4, 2, 2, 1], [8, 8, 8, 4, 4, 2, 2]
            for p in range(7):
                xsel, ysel = list(range(xo[p], w, xs[p])), list(range(yo[p], h, ys[p]))
                if xsel and ysel:
                    for y in ysel:
                        raw += b"\0" + rows_fn(xsel, y)
        else:
            for y in range(h):
                raw += b"\0" + rows_fn(list(range(w)), y)
        d = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, interlace))
        if palette is not None:
            d += chunk(b"PLTE", palette)
        if trns is not None:
            d += chunk(b"tRNS", trns)
        open(path, "wb").write(d + chunk(b"IDAT", zlib.compress(raw)) + chunk(b"IEND", b""))
    def pack_bits(vals, depth):
        out = bytearray((len(vals) * depth + 7) // 8)
        for i, v in enumerate(vals):
            b = i * depth
            out[b // 8] |= (int(v) & ((1 << depth) - 1)) << (8 - depth - b % 8)
        return bytes(out)
    w, h = 37, 23
    rgb8, rgba8 = rng.integers(0, 256, (h, w, 3), dtype=np.uint8), rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    g16, idx4, g1, g2 = rng.integers(0, 65536, (h, w), dtype=np.uint16), rng.integers(0, 16, (h, w)), rng.integers(0, 2, (h, w)), rng.integers(0, 4, (h, w))
    pal, tr = rng.integers(0, 256, 48, dtype=np.uint8).tobytes(), rng.integers(0, 256, 9, dtype=np.uint8).tobytes()
    pngs = [("i_rgb.png", lambda xs, y: rgb8[y, xs].tobytes(), w, h, 8, 2, 1, None, None), ("i_rgba.png", lambda xs, y: rgba8[y, xs].tobytes(), w, h, 8, 6, 1, None, None),
            ("i_g16.png", lambda xs, y: g16[y, xs].astype(">u2").tobytes(), w, h, 16, 0, 1, None, None),
            ("i_pal4.png", lambda xs, y: pack_bits(idx4[y, xs], 4), w, h, 4, 3, 1, pal, tr), ("n_pal4.png", lambda xs, y: pack_bits(idx4[y, xs], 4), w, h, 4, 3, 0, pal, tr),
            ("i_g1.png", lambda xs, y: pack_bits(g1[y, xs], 1), w, h, 1, 0, 1, None, None), ("i_g2.png", lambda xs, y: pack_bits(g2[y, xs], 2), w, h, 2, 0, 1, None, None),
            ("i_tiny.png", lambda xs, y: rgb8[y, xs].tobytes(), 3, 2, 8, 2, 1, None, None)]
    for name, fn, pw, ph, depth, ctype, il, pl, tn in pngs:
        path = str(tmp_path / name)
        write_png(path, fn, pw, ph, depth, ctype, il, pl, tn)
        for req in (1, 2, 3, 4):
            dump = str(tmp_path / (name + f".{req}.bin"))
            subprocess.run([REF_DUMP, "imgdump", path, str(req), dump], check=True, capture_output=True)
            out = subprocess.run([exe, path, str(req), dump], capture_output=True, text=True)
            assert out.returncode == 0 and "mismatching bytes 0 of" in out.stdout, (name, req, out.stdout)
    cmyk = str(tmp_path / "cmyk.jpg")
    img(16, 16).convert("CMYK").save(cmyk, quality=80)
    out = subprocess.run([exe, cmyk, "4", cmyk], capture_output=True, text=True)
    assert out.returncode == 2 and "3-component" in out.stdout
