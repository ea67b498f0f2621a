// Model.cpp -- OBJ / MTL ingest behind the reference's loadOBJ / addBox interface (PT_sv5_/Model.h:45-47).
//
// The reference parses with the vendored tinyobjloader and decodes textures with stb_image; neither is
// vendored here.  This is a small self-contained reader for the subset those calls exercised:
//   OBJ : v, vn, vt, f (v, v/vt, v//vn, v/vt/vn; negative indices; polygons), o, g, usemtl, mtllib
//   MTL : newmtl, Kd, Ke, map_Kd
//   textures : PNG, TGA, JPEG, BMP, PPM through host/ImageDecode.cpp (own decoders, texel-identical to the
//              reference's); undecodable files are reported and skipped (diffuseTextureID = -1) as in the reference;
//              all files the used materials name are decoded up front on all cores, ids are assigned in the reference's order
// Resulting Model, as the reference produces it (Model.cpp:138-217):
//   * a new shape starts at every `o` / `g`; one TriangleMesh per (shape, material id), shapes in file
//     order, material ids ascending inside a shape; meshes without faces are dropped;
//   * vertices are de-duplicated on the (v, vn, vt) index triple -- per MESH here.  The reference keeps that
//     map per shape, which leaves dangling indices when one shape uses two materials;
//     for single-material shapes both give identical meshes;
//   * material.color <- Kd, material.emission <- Ke, every other Disney field keeps its constructor default;
//   * faces with more than three vertices are split as a fan (tinyobjloader ear-clips: same triangles for
//     convex faces); faces without a usemtl get a default Material instead of the reference's materials[-1] read.
#include "Model.h"
#include "ImageDecode.h"

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <set>
#include <sstream>
#include <stdexcept>
#include <system_error>
#include <thread>
#include <tuple>

// ---- ownership (the containers hold raw pointers, as the reference's callers expect)
Texture::~Texture() { delete[] pixel; }

Model::~Model()
{
    for (TriangleMesh* m : meshes) delete m;
    for (Texture* t : textures) delete t;
}

namespace {

struct MtlEntry {
    std::string name;
    float3 kd = make_float3(0.6f, 0.6f, 0.6f);      // tinyobjloader's defaults when a statement is absent
    float3 ke = make_float3(0.f, 0.f, 0.f);
    std::string map_kd;
};

struct Corner { int v = -1, vt = -1, vn = -1; };
inline bool operator<(const Corner& a, const Corner& b) { return std::tie(a.v, a.vn, a.vt) < std::tie(b.v, b.vn, b.vt); }

struct Face { Corner c[3]; int material; };
struct Shape { std::string name; std::vector<Face> faces; };

std::string dir_of(const std::string& path)
{
    const size_t p = path.rfind('/');
    return p == std::string::npos ? std::string() : path.substr(0, p + 1);
}

std::string trim(const std::string& s)
{
    size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}

void parse_mtl(const std::string& path, std::vector<MtlEntry>& out, std::map<std::string, int>& by_name)
{
    std::ifstream f(path);
    if (!f) { std::cout << "Could not open material file " << path << std::endl; return; }
    std::string line;
    MtlEntry* cur = nullptr;
    while (std::getline(f, line)) {
        line = trim(line);
        if (line.empty() || line[0] == '#') continue;
        std::istringstream ss(line);
        std::string key; ss >> key;
        if (key == "newmtl") {
            MtlEntry e; std::getline(ss, e.name); e.name = trim(e.name);
            by_name[e.name] = (int)out.size();
            out.push_back(e); cur = &out.back();
        } else if (!cur) {
            continue;
        } else if (key == "Kd") {
            ss >> cur->kd.x >> cur->kd.y >> cur->kd.z;
        } else if (key == "Ke") {
            ss >> cur->ke.x >> cur->ke.y >> cur->ke.z;
        } else if (key == "map_Kd") {
            std::string rest; std::getline(ss, rest); rest = trim(rest);
            const size_t sp = rest.find_last_of(" \t");           // options (-s, -o ...) precede the file name
            cur->map_kd = sp == std::string::npos ? rest : rest.substr(sp + 1);
        }
    }
}

// "12", "12/5", "12//7", "12/5/7"; 1-based, negative = relative to the end
Corner parse_corner(const std::string& tok, int nv, int nvt, int nvn)
{
    Corner c;
    int vals[3] = {0, 0, 0}; bool has[3] = {false, false, false};
    size_t pos = 0; int field = 0;
    while (field < 3 && pos <= tok.size()) {
        size_t slash = tok.find('/', pos);
        if (slash == std::string::npos) slash = tok.size();
        if (slash > pos) { vals[field] = std::atoi(tok.substr(pos, slash - pos).c_str()); has[field] = true; }
        pos = slash + 1; ++field;
    }
    auto fix = [](int i, int n) { return i > 0 ? i - 1 : (i < 0 ? n + i : -1); };
    if (has[0]) c.v = fix(vals[0], nv);
    if (has[1]) c.vt = fix(vals[1], nvt);
    if (has[2]) c.vn = fix(vals[2], nvn);
    return c;
}

// One decoded image file; loadOBJ decodes every file its used materials name up front, on all cores (a Sponza-class scene
// names dozens of multi-megabyte PNG / TGA files: seconds of decoding when done one after the other in the mesh loop).
struct Decoded { bool ok = false; int w = 0, h = 0; std::vector<uint32_t> px; std::string why; };

std::string texture_path(const std::string& name, const std::string& model_dir)
{
    std::string file = name;
    std::replace(file.begin(), file.end(), '\\', '/');                 // "first, fix backspaces" (Model.cpp:99-102)
    return model_dir + file;
}

void decode_all(std::map<std::string, Decoded>& cache)
{
    std::vector<std::pair<const std::string, Decoded>*> jobs;
    for (auto& kv : cache) jobs.push_back(&kv);
    std::atomic<size_t> next{0};
    auto worker = [&]() {
        for (size_t i; (i = next.fetch_add(1)) < jobs.size();) {
            Decoded& d = jobs[i]->second;
            d.ok = fovpt_image::decode_file(jobs[i]->first, d.w, d.h, d.px, &d.why);
        }
    };
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const size_t nthreads = std::min<size_t>(std::min<size_t>(hw, 32), jobs.size());
    std::vector<std::thread> pool;
    try { for (size_t t = 1; t < nthreads; ++t) pool.emplace_back(worker); } catch (const std::system_error&) {}
    worker();
    for (auto& t : pool) t.join();
}

// returns the texture id in model->textures, or -1 (reference loadTexture, Model.cpp:87-136); pixels come from the cache
int load_texture(Model* model, std::map<std::string, int>& known, const std::string& name, const std::string& model_dir,
                 const std::map<std::string, Decoded>& cache)
{
    if (name.empty()) return -1;
    auto it = known.find(name);
    if (it != known.end()) return it->second;
    const std::string file = texture_path(name, model_dir);
    int id = -1;
    const Decoded& d = cache.at(file);
    const int w = d.w, h = d.h; const std::vector<uint32_t>& px = d.px; const std::string& why = d.why;
    if (d.ok) {
        Texture* t = new Texture;
        t->resolution = make_int2(w, h);
        t->pixel = new uint32_t[(size_t)w * h];
        for (int y = 0; y < h; ++y)                                    // mirror along y: row 0 becomes the bottom row (:113-124)
            std::memcpy(t->pixel + (size_t)y * w, px.data() + (size_t)(h - 1 - y) * w, (size_t)w * 4);
        id = (int)model->textures.size();
        model->textures.push_back(t);
    } else {
        std::cout << "Could not load texture from " << file << "! (" << why << ")" << std::endl;
    }
    known[name] = id;
    return id;
}

}  // namespace

Model* loadOBJ(const std::string& objFile)
{
    std::ifstream in(objFile);
    if (!in) throw std::runtime_error("Could not read OBJ model from " + objFile + " : cannot open file");
    const std::string model_dir = dir_of(objFile);

    std::vector<float3> V, VN; std::vector<float2> VT;
    std::vector<MtlEntry> materials; std::map<std::string, int> mtl_by_name;
    std::vector<Shape> shapes; shapes.emplace_back();
    int cur_mtl = -1;
    std::string line;
    while (std::getline(in, line)) {
        line = trim(line);
        if (line.empty() || line[0] == '#') continue;
        std::istringstream ss(line);
        std::string key; ss >> key;
        if (key == "v") { float3 p; ss >> p.x >> p.y >> p.z; V.push_back(p); }
        else if (key == "vn") { float3 n; ss >> n.x >> n.y >> n.z; VN.push_back(n); }
        else if (key == "vt") { float2 t = make_float2(0.f, 0.f); ss >> t.x >> t.y; VT.push_back(t); }
        else if (key == "f") {
            std::vector<Corner> poly; std::string tok;
            while (ss >> tok) poly.push_back(parse_corner(tok, (int)V.size(), (int)VT.size(), (int)VN.size()));
            for (const Corner& c : poly)
                if (c.v < 0 || c.v >= (int)V.size() || c.vt >= (int)VT.size() || c.vn >= (int)VN.size())
                    throw std::runtime_error("Could not read OBJ model from " + objFile + " : face index out of range");
            for (size_t k = 1; k + 1 < poly.size(); ++k) shapes.back().faces.push_back(Face{{poly[0], poly[k], poly[k + 1]}, cur_mtl});
        }
        else if (key == "o" || key == "g") {
            if (!shapes.back().faces.empty()) shapes.emplace_back();
            std::getline(ss, shapes.back().name); shapes.back().name = trim(shapes.back().name);
        }
        else if (key == "usemtl") {
            std::string name; std::getline(ss, name); name = trim(name);
            auto it = mtl_by_name.find(name);
            cur_mtl = it == mtl_by_name.end() ? -1 : it->second;
        }
        else if (key == "mtllib") {
            std::string name; std::getline(ss, name);
            parse_mtl(model_dir + trim(name), materials, mtl_by_name);
        }
    }
    if (shapes.back().faces.empty()) shapes.pop_back();
    std::cout << "Done loading obj file - found " << shapes.size() << " shapes with " << materials.size() << " materials" << std::endl;

    // every texture file a used material names, decoded in parallel before the meshes are assembled
    std::map<std::string, Decoded> decoded;
    for (const Shape& shape : shapes)
        for (const Face& f : shape.faces)
            if (f.material >= 0 && !materials[(size_t)f.material].map_kd.empty()) decoded[texture_path(materials[(size_t)f.material].map_kd, model_dir)];
    decode_all(decoded);

    Model* model = new Model;
    for (const Shape& shape : shapes) {
        std::set<int> ids;
        for (const Face& f : shape.faces) ids.insert(f.material);
        std::map<std::string, int> known_textures;                    // per shape, as in the reference (:175)
        for (int id : ids) {
            TriangleMesh* mesh = new TriangleMesh;
            std::map<Corner, int> known;
            auto add_vertex = [&](const Corner& c) {
                auto it = known.find(c);
                if (it != known.end()) return it->second;
                const int nid = (int)mesh->vertex.size();
                known[c] = nid;
                mesh->vertex.push_back(V[c.v]);
                if (c.vn >= 0) { while (mesh->normal.size() < mesh->vertex.size()) mesh->normal.push_back(VN[c.vn]); }
                if (c.vt >= 0) { while (mesh->texcoord.size() < mesh->vertex.size()) mesh->texcoord.push_back(VT[c.vt]); }
                if (!mesh->texcoord.empty()) mesh->texcoord.resize(mesh->vertex.size());
                if (!mesh->normal.empty()) mesh->normal.resize(mesh->vertex.size());
                return nid;
            };
            for (const Face& f : shape.faces) {
                if (f.material != id) continue;
                const int a = add_vertex(f.c[0]), b = add_vertex(f.c[1]), c = add_vertex(f.c[2]);
                mesh->index.push_back(make_uint3((unsigned)a, (unsigned)b, (unsigned)c));
            }
            mesh->diffuseTextureID = -1;
            if (id >= 0) {
                mesh->material.color = materials[id].kd;               // :190-191
                mesh->material.emission = materials[id].ke;
                mesh->diffuseTextureID = load_texture(model, known_textures, materials[id].map_kd, model_dir, decoded);
            }
            if (mesh->vertex.empty()) delete mesh; else model->meshes.push_back(mesh);
        }
    }
    std::cout << "created a total of " << model->meshes.size() << " meshes" << std::endl;
    return model;
}

void addBox(Model* model, Material& mat, const float3& pos, const float3& extend)
{
    // corner k: bit 0 -> +x, bit 1 -> +y, bit 2 -> +z
    auto corner = [&](int k) {
        return make_float3(pos.x + ((k & 1) ? extend.x : -extend.x), pos.y + ((k & 2) ? extend.y : -extend.y), pos.z + ((k & 4) ? extend.z : -extend.z));
    };
    // the reference's 12 triangles in its order and winding (front +z, back -z, left -x, right +x, top +y, bottom -y;
    // Model.cpp:232-251), written as corner numbers; 36 unshared vertices, one flat normal per face
    static const int tri[12][3] = {{4, 5, 7}, {4, 7, 6}, {0, 2, 3}, {0, 3, 1}, {0, 4, 6}, {0, 6, 2},
                                   {5, 1, 3}, {5, 3, 7}, {6, 7, 3}, {6, 3, 2}, {0, 4, 5}, {0, 5, 1}};
    static const float nrm[6][3] = {{0, 0, 1}, {0, 0, -1}, {-1, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, -1, 0}};
    TriangleMesh* mesh = new TriangleMesh;
    for (int t = 0; t < 12; ++t) {
        for (int k = 0; k < 3; ++k) {
            mesh->vertex.push_back(corner(tri[t][k]));
            mesh->normal.push_back(make_float3(nrm[t / 2][0], nrm[t / 2][1], nrm[t / 2][2]));
            mesh->texcoord.push_back(make_float2(0.f, 0.f));
        }
        mesh->index.push_back(make_uint3(3u * t, 3u * t + 1, 3u * t + 2));
    }
    mesh->material = mat;
    model->meshes.push_back(mesh);
}
