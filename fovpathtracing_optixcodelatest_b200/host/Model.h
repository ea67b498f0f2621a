// Model.h -- scene container and OBJ loading; same types and free functions as the reference's
// PT_sv5_/Model.h:10-47 (TriangleMesh, Texture, Model, loadOBJ, addBox).
#pragma once
#include <string>
#include <vector>

#include "Material.h"

struct TriangleMesh {
    std::vector<float3> vertex;
    std::vector<float3> normal;
    std::vector<float2> texcoord;
    std::vector<uint3>  index;
    Material            material;
    int                 diffuseTextureID{1};     // reference default (Model.h:19); loadOBJ always overwrites it
};

struct Texture {
    ~Texture() { delete[] pixel; }
    uint32_t* pixel{nullptr};                    // RGBA8, row 0 = bottom (loadTexture flips, Model.cpp:113-124)
    int2      resolution{-1, -1};
};

struct Model {
    ~Model()
    {
        for (auto mesh : meshes) delete mesh;
        for (auto texture : textures) delete texture;
    }
    std::vector<TriangleMesh*> meshes;
    std::vector<Texture*>      textures;
};

// One TriangleMesh per (shape, material) in shape order x ascending material id (Model.cpp:167-206);
// Kd -> material.color, Ke -> material.emission, map_Kd -> diffuseTextureID (-1 if none).
// Throws std::runtime_error when the file cannot be read (Model.cpp:160).
Model* loadOBJ(const std::string& objFile);

// axis-aligned box: 8 vertices, 12 triangles (Model.cpp:219-291)
void addBox(Model* model, Material& mat, const float3& pos, const float3& extend);
