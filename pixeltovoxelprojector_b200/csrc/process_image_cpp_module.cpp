// process_image_cpp_module.cpp -- the compiled drop-in for the reference's Python extension (boundary B-1).
//
// The reference builds `process_image_cpp` from process_image.cpp with pybind11 (setup.py:27-36, PYBIND11_MODULE at
// process_image.cpp:369-390): one function, 17 arguments (:125-143).  This is the same module surface -- same module name, function
// name, argument names, order and C++ parameter types, so pybind11 performs the same conversions and raises the same errors
// (TypeError on arity / type, forcecast copies of non-float64 arrays, ValueError from unchecked<N>() on a wrong ndim) -- with the
// body replaced by ONE call into the C ABI of libpvp_b200.so (pvp_process_image_host: copy in, K3 on the B200, copy back in place).
// It is the binding INTEGRATION.md section 1 shows, built: `make -C pixeltovoxelprojector_b200/csrc ext` puts
// process_image_cpp.<abi>.so into pixeltovoxelprojector_b200/compiled/; with that directory first on sys.path,
// `import process_image_cpp` (spacevoxelviewer.py:14) gets this module.  The pure-Python drop-in of the same name at the repo root
// additionally accepts CUDA tensors (arrays that stay on the device across calls).
//
// No compute happens here and there is no CPU fallback: without a B200 the first call raises RuntimeError.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <array>
#include <cstdlib>
#include <stdexcept>
#include <utility>
#include <vector>

#include "../../include/pvp_b200.h"

namespace py = pybind11;

namespace {

pvp_ctx *g_ctx = nullptr;

pvp_ctx *context()
{
    if (!g_ctx) {
        const char *dev = std::getenv("PVP_DEVICE");
        if (pvp_ctx_create(dev ? std::atoi(dev) : 0, nullptr, &g_ctx) != PVP_OK) throw std::runtime_error(pvp_last_error());
    }
    return g_ctx;
}

bool any_negative(const py::array &a)
{
    for (py::ssize_t k = 0; k < a.ndim(); ++k)
        if (a.strides(k) < 0) return true;
    return false;
}

// process_image.cpp:125-143
void process_image_cpp(py::array_t<double> image, std::array<double, 3> earth_position, std::array<double, 3> pointing_direction, double fov,
                       pybind11::ssize_t image_width, pybind11::ssize_t image_height, py::array_t<double> voxel_grid,
                       std::vector<std::pair<double, double>> voxel_grid_extent, double max_distance, int num_steps,
                       py::array_t<double> celestial_sphere_texture, double center_ra_rad, double center_dec_rad, double angular_width_rad,
                       double angular_height_rad, bool update_celestial_sphere, bool perform_background_subtraction)
{
    (void)image.unchecked<2>();                            // ValueError on a wrong ndim, like :146
    (void)celestial_sphere_texture.mutable_unchecked<2>();  // :147
    const bool have_grid = voxel_grid.size() > 0;          // :152 -- an empty array of any ndim skips the voxel stage
    if (have_grid) {
        (void)voxel_grid.mutable_unchecked<3>();           // :162-168
        if (voxel_grid_extent.size() < 3) throw py::index_error("voxel_grid_extent needs three (min, max) pairs");
    }
    if (image.shape(0) < image_height || image.shape(1) < image_width)   // the reference reads out of bounds here (unchecked)
        throw py::value_error("image is smaller than image_height x image_width");

    // negative strides: work on contiguous copies and assign them back (the C ABI takes non-negative element strides)
    py::array_t<double> img = any_negative(image) ? py::array_t<double>(py::array_t<double, py::array::c_style | py::array::forcecast>::ensure(image)) : image;
    py::array_t<double> tex = celestial_sphere_texture, grid = voxel_grid;
    const bool tex_copy = any_negative(tex), grid_copy = have_grid && any_negative(grid);
    if (tex_copy) tex = py::array_t<double>(py::array_t<double, py::array::c_style | py::array::forcecast>::ensure(py::array(tex.attr("copy")())));
    if (grid_copy) grid = py::array_t<double>(py::array_t<double, py::array::c_style | py::array::forcecast>::ensure(py::array(grid.attr("copy")())));

    pvp_image_args a = {};
    for (int k = 0; k < 3; ++k) { a.earth_position[k] = earth_position[k]; a.pointing_direction[k] = pointing_direction[k]; }
    a.fov = fov;
    a.image_width = image_width; a.image_height = image_height;
    a.image_stride[0] = img.strides(0) / (py::ssize_t)sizeof(double); a.image_stride[1] = img.strides(1) / (py::ssize_t)sizeof(double);
    if (have_grid) {
        for (int k = 0; k < 3; ++k) {
            a.grid_shape[k] = grid.shape(k);
            a.grid_stride[k] = grid.strides(k) / (py::ssize_t)sizeof(double);
            a.extent[2 * k] = voxel_grid_extent[k].first; a.extent[2 * k + 1] = voxel_grid_extent[k].second;
        }
    }
    a.max_distance = max_distance; a.num_steps = num_steps;
    for (int k = 0; k < 2; ++k) { a.tex_shape[k] = tex.shape(k); a.tex_stride[k] = tex.strides(k) / (py::ssize_t)sizeof(double); }
    a.center_ra_rad = center_ra_rad; a.center_dec_rad = center_dec_rad; a.angular_width_rad = angular_width_rad; a.angular_height_rad = angular_height_rad;
    a.update_celestial_sphere = update_celestial_sphere; a.perform_background_subtraction = perform_background_subtraction;
    a.accum_mode = PVP_ACCUM_F64; a.frac_bits = 0;

    int rc;
    {
        py::gil_scoped_release nogil;   // (the reference holds the GIL for the whole call; nothing here touches Python objects)
        rc = pvp_process_image_host(context(), img.data(), have_grid ? grid.mutable_data() : nullptr, tex.mutable_data(), &a, nullptr);
    }
    if (rc != PVP_OK) throw std::runtime_error(pvp_last_error());
    if (tex_copy) celestial_sphere_texture.attr("__setitem__")(py::ellipsis(), tex);
    if (grid_copy) voxel_grid.attr("__setitem__")(py::ellipsis(), grid);
}

}  // namespace

// process_image.cpp:369-390
PYBIND11_MODULE(process_image_cpp, m)
{
    m.doc() = "B200 implementation of the process_image function (drop-in for the reference's process_image_cpp extension)";
    m.def("process_image_cpp", &process_image_cpp, "Process image and update voxel grid on the B200", py::arg("image"), py::arg("earth_position"),
          py::arg("pointing_direction"), py::arg("fov"), py::arg("image_width"), py::arg("image_height"), py::arg("voxel_grid"),
          py::arg("voxel_grid_extent"), py::arg("max_distance"), py::arg("num_steps"), py::arg("celestial_sphere_texture"),
          py::arg("center_ra_rad"), py::arg("center_dec_rad"), py::arg("angular_width_rad"), py::arg("angular_height_rad"),
          py::arg("update_celestial_sphere"), py::arg("perform_background_subtraction"));
}
