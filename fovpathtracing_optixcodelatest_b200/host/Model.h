// Model.h -- host-side scene container of the drop-in: the type and member NAMES are the reference's
// (PT_sv5_/Model.h:10-47: TriangleMesh, Texture, Model, loadOBJ, addBox) because main.cpp and
// SampleRenderer address them by name; everything else (layout order, ownership code in Model.cpp,
// the OBJ/MTL reader) is this repository's.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "Material.h"

// One (shape, material) group of an OBJ file -- what becomes one mesh of fovpt_scene_add_mesh.
struct TriangleMesh {
    Material material;                  // Disney parameters; loadOBJ sets color <- Kd and emission <- Ke only
    int diffuseTextureID = 1;           // index into Model::textures or -1; 1 is the reference's odd default
                                        // (Model.h:19), loadOBJ always overwrites it
    std::vector<uint3>  index;          // triangles, indices into `vertex`
    std::vector<float3> vertex;
    std::vector<float3> normal;         // carried for API compatibility; shading uses geometric normals
    std::vector<float2> texcoord;       // empty or one per vertex
};

// RGBA8 image, row 0 at the bottom (the loader flips, as reference Model.cpp:113-124 does).
struct Texture {
    ~Texture();
    int2      resolution = {-1, -1};
    uint32_t* pixel = nullptr;          // new[]-allocated, owned
};

// Owns its meshes and textures (raw pointers, as the reference's callers expect).
struct Model {
    ~Model();
    std::vector<Texture*>      textures;
    std::vector<TriangleMesh*> meshes;
};

// Meshes come out in shape order x ascending material id (reference Model.cpp:167-206); map_Kd ->
// diffuseTextureID.  Throws std::runtime_error if the file cannot be read (Model.cpp:160).
Model* loadOBJ(const std::string& objFile);

// Appends an axis-aligned box (8 vertices, 12 triangles in the reference's order, Model.cpp:219-291).
void addBox(Model* model, Material& mat, const float3& pos, const float3& extend);
