#pragma once
#include <cuda_runtime.h>
#include <cstddef>

namespace vgh {
struct FlameModel;
int flame_model_create(const float* v_template, const float* shapedirs, const float* posedirs, const float* j_regressor,
                       const float* lbs_weights, FlameModel** out, char* err, size_t errlen);
void flame_model_destroy(FlameModel* m);
// params [n,413] device; n_dev optional device-side count; ns/ne = live shape/expression coefficients;
// xform optional [n,3] (pad_x, pad_y, img_scale); verts/rot optional outputs; proj [n,5023,3].
int flame_decode_launch(const FlameModel* m, const float* params, int n, const int* n_dev, int ns, int ne,
                        const float* xform, float* verts, float* rot, float* proj, cudaStream_t stream, char* err,
                        size_t errlen);
}  // namespace vgh
