// Storage shared by the scene-file loaders (host_pbrt.cpp, host_mitsuba.cpp): what a loader collects and the
// yk_host_scene_desc view over it. Both loaders hand out the same handle type (yk_pbrt_scene), read with yk_pbrt_view.
#pragma once
#include <cstring>
#include <string>
#include <vector>

#include "host_math.h"
#include "yuki_gpu.h"

struct YkMeshStore {
    ykh::xform o2w;
    std::vector<float> points, normals, uvs;
    std::vector<uint32_t> indices;
    int32_t material;
};

struct yk_pbrt_scene {
    std::vector<YkMeshStore> meshes;
    std::vector<yk_mesh_desc> mesh_descs;
    std::vector<yk_sphere_desc> spheres;
    std::vector<int32_t> objects;  // file order: mesh index, or -1 - sphere index
    std::vector<yk_texture_desc> textures;
    std::vector<std::vector<float>> texel_storage;
    std::vector<yk_material_desc> materials;
    std::vector<yk_light_desc> lights;
    yk_pbrt_result result{};

    // Points result.scene at the collected storage (call once, after the last push_back).
    void finish(uint32_t max_shapes_in_node, uint32_t split_method) {
        mesh_descs.clear();
        for (const YkMeshStore& m : meshes) {
            yk_mesh_desc d{};
            std::memcpy(d.object_to_world.m, m.o2w.m.e, 64);
            std::memcpy(d.object_to_world.m_inv, m.o2w.inv.e, 64);
            d.n_points = (uint32_t)(m.points.size() / 3);
            d.n_indices = (uint32_t)m.indices.size();
            d.points = m.points.data();
            d.normals = m.normals.size() == m.points.size() && !m.normals.empty() ? m.normals.data() : nullptr;
            d.uvs = m.uvs.size() / 2 == m.points.size() / 3 && !m.uvs.empty() ? m.uvs.data() : nullptr;
            d.indices = m.indices.data();
            d.material = m.material;
            d.area_light = -1;  // neither file format's loader creates area lights (pbrt/mod.rs:502)
            mesh_descs.push_back(d);
        }
        yk_host_scene_desc& hd = result.scene;
        hd.n_meshes = (uint32_t)mesh_descs.size();
        hd.meshes = mesh_descs.data();
        hd.n_textures = (uint32_t)textures.size();
        hd.textures = textures.data();
        hd.n_materials = (uint32_t)materials.size();
        hd.materials = materials.data();
        hd.n_lights = (uint32_t)lights.size();
        hd.lights = lights.data();
        hd.max_shapes_in_node = max_shapes_in_node ? max_shapes_in_node : 1u;
        hd.split_method = split_method;
        hd.n_spheres = (uint32_t)spheres.size();
        hd.spheres = spheres.data();
        hd.n_objects = (uint32_t)objects.size();
        hd.objects = objects.data();
    }
};
