// pslam_plugin_capi.cpp -- extern "C" view of the plugin mirror (include/pslam_plugin.h).  Exceptions thrown by
// the C++ modules (the reference's std::runtime_error texts) are caught at this boundary.
#include <cstring>
#include <map>
#include <memory>

#include "../../include/pslam_plugin.h"
#include "pslam_plugin.hpp"

using namespace pslam_host;

struct psp_manager {
  ConfigurableManager manager;
};

namespace {

thread_local std::string g_error;

// clouds / result buffers handed over through the C interface, kept alive per module
struct Storage {
  PointIntensityDescriptorCloud fixed{3}, moving{3}, meas{4};
  CorrespondenceVector correspondences;
};
std::map<const Configurable*, std::unique_ptr<Storage>> g_storage;

Storage& storage_of(const Configurable* c) {
  auto& s = g_storage[c];
  if (!s) s.reset(new Storage());
  return *s;
}

Configurable* mod(psp_module* m) { return reinterpret_cast<Configurable*>(m); }
psp_module* handle(Configurable* c) { return reinterpret_cast<psp_module*>(c); }

template <typename T>
T* as(psp_module* m, const char* what) {
  T* t = dynamic_cast<T*>(mod(m));
  if (!t) throw std::runtime_error(std::string("module is not a ") + what);
  return t;
}

template <typename F>
int guard(F&& f) {
  try {
    g_error.clear();
    return f();
  } catch (const std::exception& e) {
    g_error = e.what();
    return -1;
  } catch (...) {
    g_error = "unknown exception";
    return -1;
  }
}

void set_cloud(PointIntensityDescriptorCloud& c, int n, int dim, const float* coords, const uint8_t* desc) {
  c.dim = dim;
  c.number_of_optimizations.clear();
  c.resize(n);
  if (n > 0) {
    std::memcpy(c.coordinates.data(), coords, sizeof(float) * (size_t) n * dim);
    std::memcpy(c.descriptor.data(), desc, 32 * (size_t) n);
  }
  std::fill(c.intensity.begin(), c.intensity.end(), 0.f);
}

int copy_out(const CorrespondenceVector& v, int capacity, int* f, int* m, float* r) {
  const int n = (int) v.size();
  for (int i = 0; i < n && i < capacity; ++i) {
    if (f) f[i] = v[i].fixed_idx;
    if (m) m[i] = v[i].moving_idx;
    if (r) r[i] = v[i].response;
  }
  return n;
}

int copy_cloud(const PointIntensityDescriptorCloud& c, int capacity, float* coords, float* intensity, uint8_t* desc) {
  const int n = (int) c.size();
  const int m = n < capacity ? n : capacity;
  if (m > 0) {
    if (coords) std::memcpy(coords, c.coordinates.data(), sizeof(float) * (size_t) m * c.dim);
    if (intensity) std::memcpy(intensity, c.intensity.data(), sizeof(float) * m);
    if (desc) std::memcpy(desc, c.descriptor.data(), 32 * (size_t) m);
  }
  return n;
}

}  // namespace

extern "C" {

const char* psp_last_error(void) { return g_error.c_str(); }

int psp_set_device(int device) {
  return guard([&] {
    PslamDevice::setDevice(device);
    return 0;
  });
}

// per-kernel device timing of the process-wide context (pslam_profile_*): tools/aligner_probe.py, bench.py
int psp_profile_enable(int enable) {
  return guard([&] { return pslam_profile_enable(PslamDevice::context(), enable); });
}
int psp_profile_read(int capacity, char* names, int name_len, double* total_ms, long long* launches) {
  return guard([&] { return pslam_profile_read(PslamDevice::context(), capacity, names, name_len, total_ms, launches); });
}

psp_manager* psp_manager_create(void) {
  registerTypes();
  return new psp_manager();
}

void psp_manager_destroy(psp_manager* m) {
  if (!m) return;
  for (const ConfigurablePtr& c : m->manager.instances()) g_storage.erase(c.get());
  delete m;
}

int psp_manager_read(psp_manager* m, const char* conf_path) {
  return guard([&] {
    m->manager.read(conf_path);
    return (int) m->manager.instances().size();
  });
}

int psp_manager_read_string(psp_manager* m, const char* conf_text) {
  return guard([&] {
    m->manager.readString(conf_text);
    return (int) m->manager.instances().size();
  });
}

int psp_manager_write(psp_manager* m, const char* conf_path, int n_names, const char* const* names) {
  return guard([&] {
    std::vector<ConfigurablePtr> roots;
    for (int i = 0; i < n_names; ++i) {
      ConfigurablePtr c = m->manager.getByName(names[i]);
      if (!c) throw std::runtime_error(std::string("no module named '") + names[i] + "'");
      roots.push_back(c);
    }
    m->manager.write(conf_path, roots);
    return 0;
  });
}

int psp_manager_count(psp_manager* m) { return (int) m->manager.instances().size(); }

psp_module* psp_manager_at(psp_manager* m, int index) {
  if (index < 0 || index >= (int) m->manager.instances().size()) return nullptr;
  return handle(m->manager.instances()[index].get());
}

psp_module* psp_manager_get_by_name(psp_manager* m, const char* name) { return handle(m->manager.getByName(name).get()); }

psp_module* psp_manager_create_module(psp_manager* m, const char* class_name, const char* name) {
  registerTypes();
  return handle(m->manager.create(class_name, name ? name : "").get());
}

int psp_class_is_registered(const char* class_name) {
  registerTypes();
  return ClassRegistry::instance().has(class_name) ? 1 : 0;
}

const char* psp_module_class_name(psp_module* c) { return mod(c)->className().c_str(); }
const char* psp_module_name(psp_module* c) { return mod(c)->name().c_str(); }
int psp_module_is_generic(psp_module* c) { return dynamic_cast<GenericConfigurable*>(mod(c)) ? 1 : 0; }
int psp_module_has_param(psp_module* c, const char* param) {
  return (mod(c)->property(param) || mod(c)->extraValues().count(param)) ? 1 : 0;
}

int psp_module_set_number(psp_module* c, const char* param, double value) {
  return guard([&] {
    PropertyBase* p = mod(c)->property(param);
    if (!p) {
      mod(c)->extraValues()[param] = ConfValue::num(value);
      return 0;
    }
    p->fromConf(ConfValue::num(value), Resolver());
    return 0;
  });
}

int psp_module_get_number(psp_module* c, const char* param, double* value) {
  return guard([&] {
    ConfValue v;
    if (PropertyBase* p = mod(c)->property(param)) {
      v = p->toConf(IdOf());
    } else {
      auto it = mod(c)->extraValues().find(param);
      if (it == mod(c)->extraValues().end()) throw std::runtime_error(std::string("no parameter '") + param + "'");
      v = it->second;
    }
    if (v.kind != ConfValue::Number) throw std::runtime_error(std::string("parameter '") + param + "' is not a number");
    *value = v.number;
    return 0;
  });
}

int psp_module_set_string(psp_module* c, const char* param, const char* value) {
  return guard([&] {
    PropertyBase* p = mod(c)->property(param);
    if (!p) {
      mod(c)->extraValues()[param] = ConfValue::str(value);
      return 0;
    }
    p->fromConf(ConfValue::str(value), Resolver());
    return 0;
  });
}

const char* psp_module_get_string(psp_module* c, const char* param) {
  static thread_local std::string out;
  out.clear();
  if (PropertyBase* p = mod(c)->property(param)) {
    ConfValue v = p->toConf(IdOf());
    if (v.kind == ConfValue::String) out = v.text;
  } else {
    auto it = mod(c)->extraValues().find(param);
    if (it != mod(c)->extraValues().end() && it->second.kind == ConfValue::String) out = it->second.text;
  }
  return out.c_str();
}

int psp_module_set_numbers(psp_module* c, const char* param, int n, const double* values) {
  return guard([&] {
    ConfValue v;
    v.kind = ConfValue::Array;
    for (int i = 0; i < n; ++i) v.items.push_back(ConfValue::num(values[i]));
    PropertyBase* p = mod(c)->property(param);
    if (!p) {
      mod(c)->extraValues()[param] = v;
      return 0;
    }
    p->fromConf(v, Resolver());
    return 0;
  });
}

int psp_module_get_numbers(psp_module* c, const char* param, int capacity, double* values) {
  return guard([&] {
    ConfValue v;
    if (PropertyBase* p = mod(c)->property(param)) {
      v = p->toConf([](const Configurable*) { return -1; });
    } else {
      auto it = mod(c)->extraValues().find(param);
      if (it == mod(c)->extraValues().end()) throw std::runtime_error(std::string("no parameter '") + param + "'");
      v = it->second;
    }
    if (v.kind != ConfValue::Array) throw std::runtime_error(std::string("parameter '") + param + "' is not an array");
    int n = 0;
    for (const ConfValue& i : v.items) {
      if (i.kind != ConfValue::Number) continue;
      if (n < capacity) values[n] = i.number;
      ++n;
    }
    return n;
  });
}

int psp_module_set_link(psp_module* c, const char* param, psp_module* target) {
  return guard([&] {
    auto* p = dynamic_cast<PropertyConfigurableBase*>(mod(c)->property(param));
    if (!p) throw std::runtime_error(std::string("no link parameter '") + param + "'");
    p->setPointer(target ? mod(target)->shared_from_this() : nullptr);
    return 0;
  });
}

psp_module* psp_module_get_link(psp_module* c, const char* param) {
  auto* p = dynamic_cast<PropertyConfigurableBase*>(mod(c)->property(param));
  if (!p) return nullptr;
  return handle(p->pointer().get());
}

int psp_extractor_paint_tracking_mask(psp_module* extractor, int rows, int cols, int n, int dim, const float* coords, int radius,
                                      uint8_t* mask) {
  return guard([&] {
    auto* ex = as<IntensityFeatureExtractorSelectiveCUDA>(extractor, "IntensityFeatureExtractorSelective");
    PointIntensityDescriptorCloud proj(dim);
    proj.resize((size_t) n);
    if (n > 0) std::memcpy(proj.coordinates.data(), coords, sizeof(float) * (size_t) n * dim);
    std::vector<uint8_t> m;
    ex->paintTrackingMask(rows, cols, proj, radius, m);
    std::memcpy(mask, m.data(), m.size());
    return 0;
  });
}

int psp_extractor_compute(psp_module* extractor, const uint8_t* image, int rows, int cols, int stride,
                          const uint8_t* mask, int capacity, float* xy, float* intensity, uint8_t* desc) {
  return guard([&] {
    auto* ex = as<IntensityFeatureExtractorBaseCUDA>(extractor, "IntensityFeatureExtractor");
    PointIntensityDescriptorCloud cloud(ex->pointDim());
    ex->setFeatures(&cloud);
    ImageView m;
    if (mask) m = ImageView{mask, rows, cols, cols};
    ex->setKeypointDetectionMask(m);
    ex->compute(ImageView{image, rows, cols, stride});
    ex->setFeatures(nullptr);
    const int n = (int) cloud.size();
    for (int i = 0; i < n && i < capacity; ++i) {
      if (xy) {
        xy[2 * i] = cloud.point(i)[0];
        xy[2 * i + 1] = cloud.point(i)[1];
      }
    }
    copy_cloud(cloud, capacity, nullptr, intensity, desc);
    return n;
  });
}

int psp_extractor_set_projections(psp_module* extractor, int n, int dim, const float* coords, int radius) {
  return guard([&] {
    auto* ex = as<IntensityFeatureExtractorBaseCUDA>(extractor, "IntensityFeatureExtractor");
    Storage& st = storage_of(mod(extractor));
    st.moving.dim = dim;
    st.moving.number_of_optimizations.clear();
    st.moving.resize(n);
    if (n > 0) std::memcpy(st.moving.coordinates.data(), coords, sizeof(float) * (size_t) n * dim);
    ex->setProjections(n > 0 ? &st.moving : nullptr, (size_t) radius);
    return 0;
  });
}

int psp_extractor_number_of_tracking_keypoints(psp_module* extractor) {
  return guard([&] {
    auto* ex = as<IntensityFeatureExtractorSelectiveCUDA>(extractor, "IntensityFeatureExtractorSelective");
    return (int) ex->numberOfTrackingKeypoints();
  });
}

int psp_stereo_adaptor_compute(psp_module* adaptor, const uint8_t* left, const uint8_t* right, int rows, int cols,
                               int stride, int capacity, float* uvuv, float* intensity, uint8_t* desc, int* status) {
  return guard([&] {
    auto* ad = as<RawDataPreprocessorStereoProjectiveCUDA>(adaptor, "RawDataPreprocessorStereoProjective");
    Storage& st = storage_of(mod(adaptor));
    ad->setMeas(&st.meas);
    struct StatusOut {
      RawDataPreprocessorBase* a;
      int* s;
      ~StatusOut() {
        if (s) *s = (int) a->status();
      }
    } so{ad, status};
    ad->setRawData(ImageView{left, rows, cols, stride}, ImageView{right, rows, cols, stride});
    ad->compute();
    return copy_cloud(st.meas, capacity, uvuv, intensity, desc);
  });
}

int psp_mono_depth_adaptor_compute(psp_module* adaptor, const uint8_t* image, int rows, int cols, int stride,
                                   const void* depth, int depth_type, int depth_rows, int depth_cols,
                                   int depth_stride_elements, int capacity, float* uvz, float* intensity,
                                   uint8_t* desc, int* status) {
  return guard([&] {
    auto* ad = as<RawDataPreprocessorMonocularDepthCUDA>(adaptor, "RawDataPreprocessorMonocularDepth");
    Storage& st = storage_of(mod(adaptor));
    ad->setMeas(&st.meas);
    struct StatusOut {
      RawDataPreprocessorBase* a;
      int* s;
      ~StatusOut() {
        if (s) *s = (int) a->status();
      }
    } so{ad, status};
    ad->setRawData(ImageView{image, rows, cols, stride}, DepthView{depth, depth_rows, depth_cols, depth_stride_elements, depth_type});
    ad->compute();
    return copy_cloud(st.meas, capacity, uvz, intensity, desc);
  });
}

int psp_finder_set_fixed(psp_module* finder, int n, int dim, const float* coords, const uint8_t* desc) {
  return guard([&] {
    auto* f = as<CorrespondenceFinderBase>(finder, "CorrespondenceFinder");
    Storage& st = storage_of(mod(finder));
    set_cloud(st.fixed, n, dim, coords, desc);
    f->setFixed(&st.fixed);
    f->setCorrespondences(&st.correspondences);
    return 0;
  });
}

int psp_finder_set_moving(psp_module* finder, int n, int dim, const float* coords, const uint8_t* desc) {
  return guard([&] {
    auto* f = as<CorrespondenceFinderBase>(finder, "CorrespondenceFinder");
    Storage& st = storage_of(mod(finder));
    set_cloud(st.moving, n, dim, coords, desc);
    f->setMoving(&st.moving);
    f->setCorrespondences(&st.correspondences);
    return 0;
  });
}

int psp_finder_set_local_map_in_sensor(psp_module* finder, const float* pose12) {
  return guard([&] {
    auto* f = as<CorrespondenceFinderBase>(finder, "CorrespondenceFinder");
    Isometry3f T;
    std::memcpy(T.m, pose12, sizeof(T.m));
    f->setLocalMapInSensor(T);
    return 0;
  });
}

int psp_finder_compute(psp_module* finder, int capacity, int* fixed_idx, int* moving_idx, float* response) {
  return guard([&] {
    auto* f = as<CorrespondenceFinderBase>(finder, "CorrespondenceFinder");
    Storage& st = storage_of(mod(finder));
    f->compute();
    return copy_out(st.correspondences, capacity, fixed_idx, moving_idx, response);
  });
}

int psp_projective_finder_state(psp_module* finder, int* search_radius_pixels, float* descriptor_distance,
                                int* current_iteration, int* has_converged, int* number_of_searches) {
  return guard([&] {
    auto* f = as<CorrespondenceFinderProjectiveCUDA>(finder, "CorrespondenceFinderProjective");
    if (search_radius_pixels) *search_radius_pixels = (int) f->searchRadiusPixels();
    if (descriptor_distance) *descriptor_distance = f->descriptorDistance();
    if (current_iteration) *current_iteration = (int) f->currentIteration();
    if (has_converged) *has_converged = f->hasConverged() ? 1 : 0;
    if (number_of_searches) *number_of_searches = f->numberOfSearches();
    return 0;
  });
}

int psp_projective_finder_set_state(psp_module* finder, int search_radius_pixels, float descriptor_distance) {
  return guard([&] {
    auto* f = as<CorrespondenceFinderProjectiveCUDA>(finder, "CorrespondenceFinderProjective");
    f->setSearchradiusPixels((size_t) search_radius_pixels);
    f->setDescriptorDistance(descriptor_distance);
    return 0;
  });
}

int psp_projector_set_camera_matrix(psp_module* projector, const float* K9) {
  return guard([&] {
    auto* p = as<ProjectorPinhole>(projector, "PointIntensityDescriptor3fProjectorPinhole");
    std::array<float, 9> K;
    std::memcpy(K.data(), K9, sizeof(float) * 9);
    p->setCameraMatrix(K);
    return 0;
  });
}

int psp_aligner_set_fixed(psp_module* aligner, int n, int dim, const float* coords, const uint8_t* desc) {
  return guard([&] {
    auto* a = as<MultiAligner3DQRCUDA>(aligner, "MultiAligner3DQR");
    Storage& st = storage_of(mod(aligner));
    set_cloud(st.fixed, n, dim, coords, desc);
    a->setFixed(&st.fixed);
    return 0;
  });
}

int psp_aligner_set_moving(psp_module* aligner, int n, const float* xyz, const uint8_t* desc, const int* n_opt) {
  return guard([&] {
    auto* a = as<MultiAligner3DQRCUDA>(aligner, "MultiAligner3DQR");
    Storage& st = storage_of(mod(aligner));
    set_cloud(st.moving, n, 3, xyz, desc);
    if (n_opt) st.moving.number_of_optimizations.assign(n_opt, n_opt + n);
    a->setMoving(&st.moving);
    return 0;
  });
}

int psp_aligner_set_moving_in_fixed(psp_module* aligner, const float* pose12) {
  return guard([&] {
    auto* a = as<MultiAligner3DQRCUDA>(aligner, "MultiAligner3DQR");
    Isometry3f T;
    std::memcpy(T.m, pose12, sizeof(T.m));
    a->setMovingInFixed(T);
    return 0;
  });
}

int psp_aligner_set_left_camera_in_right(psp_module* aligner, const float* t3) {
  return guard([&] {
    auto* a = as<MultiAligner3DQRCUDA>(aligner, "MultiAligner3DQR");
    AlignerSliceProcessorProjectiveCUDA* s = a->projectiveSlice();
    if (!s) throw std::runtime_error("MultiAligner|ERROR: no projective slice processor configured");
    s->setLeftCameraInRight(t3);
    return 0;
  });
}

int psp_aligner_compute(psp_module* aligner, double* moving_in_fixed12, int* iterations, int* num_correspondences,
                        int* num_inliers, double* chi) {
  return guard([&] {
    auto* a = as<MultiAligner3DQRCUDA>(aligner, "MultiAligner3DQR");
    a->compute();
    if (moving_in_fixed12) std::memcpy(moving_in_fixed12, a->movingInFixed().data(), sizeof(double) * 12);
    const auto& st = a->iterationStats();
    if (iterations) *iterations = (int) st.size();
    if (!st.empty()) {
      if (num_correspondences) *num_correspondences = st.back().num_correspondences;
      if (num_inliers) *num_inliers = st.back().num_inliers;
      if (chi) *chi = st.back().chi;
    }
    return (int) a->status();
  });
}

int psp_aligner_set_slice_processor(psp_module* aligner, int index, psp_module* slice) {
  return guard([&] {
    auto* a = as<MultiAligner3DQRCUDA>(aligner, "MultiAligner3DQR");
    if (!slice) throw std::runtime_error("MultiAligner|ERROR: null slice processor");
    a->param_slice_processors.setValue((size_t) index, mod(slice)->shared_from_this());
    return 0;
  });
}

int psp_aligner_num_slice_processors(psp_module* aligner) {
  return guard([&] { return (int) as<MultiAligner3DQRCUDA>(aligner, "MultiAligner3DQR")->param_slice_processors.size(); });
}

int psp_aligner_set_trajectory_chunk(psp_module* aligner, int n, const float* poses12) {
  return guard([&] {
    auto* a = as<MultiAligner3DQRCUDA>(aligner, "MultiAligner3DQR");
    AlignerSliceMotionModel3DCUDA* s = a->motionModelSlice();
    if (!s) throw std::runtime_error("MultiAligner|ERROR: no motion model slice configured");
    std::vector<Isometry3f> chunk((size_t) n);
    for (int i = 0; i < n; ++i) std::memcpy(chunk[i].m, poses12 + 12 * (size_t) i, sizeof(float) * 12);
    s->setTrajectoryChunk(chunk);
    return 0;
  });
}

int psp_aligner_set_prior_information(psp_module* aligner, const double* information36) {
  return guard([&] {
    auto* a = as<MultiAligner3DQRCUDA>(aligner, "MultiAligner3DQR");
    AlignerSliceMotionModel3DCUDA* s = a->motionModelSlice();
    if (!s) throw std::runtime_error("MultiAligner|ERROR: no motion model slice configured");
    s->setInformationMatrix(information36);
    return 0;
  });
}

int psp_aligner_inlier_run_stats(psp_module* aligner, int capacity, double* rows4) {
  return guard([&] {
    auto* a = as<MultiAligner3DQRCUDA>(aligner, "MultiAligner3DQR");
    const auto& st = a->inlierRunStats();
    for (int i = 0; i < (int) st.size() && i < capacity; ++i) {
      rows4[4 * i] = st[i].num_correspondences;
      rows4[4 * i + 1] = st[i].num_inliers;
      rows4[4 * i + 2] = st[i].num_outliers;
      rows4[4 * i + 3] = st[i].chi;
    }
    return (int) st.size();
  });
}

int psp_aligner_iteration_stats(psp_module* aligner, int capacity, double* rows4) {
  return guard([&] {
    auto* a = as<MultiAligner3DQRCUDA>(aligner, "MultiAligner3DQR");
    const auto& st = a->iterationStats();
    for (int i = 0; i < (int) st.size() && i < capacity; ++i) {
      rows4[4 * i] = st[i].num_correspondences;
      rows4[4 * i + 1] = st[i].num_inliers;
      rows4[4 * i + 2] = st[i].num_outliers;
      rows4[4 * i + 3] = st[i].chi;
    }
    return (int) st.size();
  });
}

int psp_aligner_correspondences(psp_module* aligner, int capacity, int* fixed_idx, int* moving_idx, float* response) {
  return guard([&] {
    auto* a = as<MultiAligner3DQRCUDA>(aligner, "MultiAligner3DQR");
    AlignerSliceProcessorProjectiveCUDA* s = a->projectiveSlice();
    if (!s) throw std::runtime_error("MultiAligner|ERROR: no projective slice processor configured");
    return copy_out(s->correspondences(), capacity, fixed_idx, moving_idx, response);
  });
}

// ---- SceneClipperProjective3D ----------------------------------------------------------------------------------
int psp_clipper_set_full_scene(psp_module* clipper, int n, const float* xyz, const float* intensity, const uint8_t* desc) {
  return guard([&] {
    auto* c = as<SceneClipperProjective3DCUDA>(clipper, "SceneClipperProjective3D");
    Storage& st = storage_of(mod(clipper));
    st.moving.dim = 3;
    st.moving.number_of_optimizations.clear();
    st.moving.resize(n);
    if (n > 0) {
      std::memcpy(st.moving.coordinates.data(), xyz, sizeof(float) * 3 * (size_t) n);
      if (intensity) std::memcpy(st.moving.intensity.data(), intensity, sizeof(float) * (size_t) n);
      else std::fill(st.moving.intensity.begin(), st.moving.intensity.end(), 0.f);
      std::memcpy(st.moving.descriptor.data(), desc, 32 * (size_t) n);
    }
    c->setFullScene(&st.moving);
    return 0;
  });
}

int psp_clipper_set_robot_in_local_map(psp_module* clipper, const float* pose12) {
  return guard([&] {
    Isometry3f T;
    std::memcpy(T.m, pose12, sizeof(T.m));
    as<SceneClipperProjective3DCUDA>(clipper, "SceneClipperProjective3D")->setRobotInLocalMap(T);
    return 0;
  });
}

int psp_clipper_set_sensor_in_robot(psp_module* clipper, const float* pose12) {
  return guard([&] {
    Isometry3f T;
    std::memcpy(T.m, pose12, sizeof(T.m));
    as<SceneClipperProjective3DCUDA>(clipper, "SceneClipperProjective3D")->setSensorInRobot(T);
    return 0;
  });
}

int psp_clipper_compute(psp_module* clipper, int capacity, float* xyz, float* uvz, int* global_index, uint8_t* desc,
                        int* status) {
  return guard([&] {
    auto* c = as<SceneClipperProjective3DCUDA>(clipper, "SceneClipperProjective3D");
    Storage& st = storage_of(mod(clipper));
    st.fixed.dim = 3;
    c->setClippedSceneInRobot(&st.fixed);
    struct StatusOut {
      SceneClipperProjective3DCUDA* c;
      int* s;
      ~StatusOut() {
        if (s) *s = (int) c->status();
      }
    } so{c, status};
    c->compute();
    if (c->status() != SceneClipperProjective3DCUDA::Successful) return 0;
    const int n = copy_cloud(st.fixed, capacity, xyz, nullptr, desc);
    const int m = n < capacity ? n : capacity;
    if (m > 0) {
      if (uvz) std::memcpy(uvz, c->projections().data(), sizeof(float) * 3 * (size_t) m);
      if (global_index) std::memcpy(global_index, c->globalIndices().data(), sizeof(int) * (size_t) m);
    }
    return n;
  });
}

// ---- point EKFs + LandmarkEstimatorEKF --------------------------------------------------------------------------
int psp_point_filter_set_camera(psp_module* filter, const float* K9, double baseline_x_pixels, double baseline_y_pixels) {
  return guard([&] {
    auto* f = as<PointEKFCUDA>(filter, "PointEKF");
    std::array<float, 9> K;
    std::memcpy(K.data(), K9, sizeof(float) * 9);
    f->setCameraMatrix(K);
    f->setBaseline(baseline_x_pixels, baseline_y_pixels);
    return 0;
  });
}

int psp_landmark_estimator_set_transforms(psp_module* estimator, const float* measurement_in_world12, const float* measurement_in_scene12) {
  return guard([&] {
    Isometry3f W, S;
    std::memcpy(W.m, measurement_in_world12, sizeof(W.m));
    std::memcpy(S.m, measurement_in_scene12, sizeof(S.m));
    if (auto* wm = dynamic_cast<LandmarkEstimatorWeightedMeanCUDA*>(mod(estimator))) wm->setTransforms(W, S);
    else if (auto* sm = dynamic_cast<LandmarkEstimatorPoseBasedSmootherCUDA*>(mod(estimator))) sm->setTransforms(W, S);
    else as<LandmarkEstimatorEKFCUDA>(estimator, "LandmarkEstimatorEKF")->setTransforms(W, S);
    return 0;
  });
}

int psp_landmark_estimator_compute_batch(psp_module* estimator, int n, float* state_world, float* covariance,
                                         const float* measurements, float* coords_in_local_map, uint8_t* inlier) {
  return guard([&] {
    return as<LandmarkEstimatorEKFCUDA>(estimator, "LandmarkEstimatorEKF")
      ->computeBatch(n, state_world, covariance, measurements, coords_in_local_map, inlier);
  });
}

int psp_landmark_estimator_weighted_mean_batch(psp_module* estimator, int n, float* state_world, const int* number_of_optimizations,
                                               const float* landmark_in_sensor, float* coords_in_local_map, uint8_t* inlier) {
  return guard([&] {
    return as<LandmarkEstimatorWeightedMeanCUDA>(estimator, "LandmarkEstimatorWeightedMean")
      ->computeBatch(n, state_world, number_of_optimizations, landmark_in_sensor, coords_in_local_map, inlier);
  });
}

int psp_merger_select_updates(psp_module* merger, const float* measurements, int dim, int n_meas, const int* corr_moving,
                              const float* corr_response, int n_corr, uint8_t* selected) {
  return guard([&] {
    return as<MergerProjectiveCUDA>(merger, "MergerProjective")
      ->selectUpdates(measurements, dim, n_meas, corr_moving, corr_response, n_corr, selected);
  });
}

int psp_merger_plan(psp_module* merger, const float* measurements, int dim, int n_meas, const int* corr_moving, const float* corr_response,
                    int n_corr, uint8_t* selected, int* winners, int* n_winners) {
  return guard([&] {
    return as<MergerProjectiveCUDA>(merger, "MergerProjective")
      ->plan(measurements, dim, n_meas, corr_moving, corr_response, n_corr, selected, winners, n_winners);
  });
}

int psp_merger_wants_additions(psp_module* merger, int merged, int n_meas, int n_corr) {
  return guard([&] { return as<MergerProjectiveCUDA>(merger, "MergerProjective")->wantsAdditions(merged, n_meas, n_corr) ? 1 : 0; });
}

int psp_merger_select_additions(psp_module* merger, const float* measurements, int dim, int n_meas, int* winners) {
  return guard([&] { return as<MergerProjectiveCUDA>(merger, "MergerProjective")->selectAdditions(measurements, dim, n_meas, winners); });
}

int psp_landmark_smoother_set_camera_matrix(psp_module* estimator, const float* K9) {
  return guard([&] {
    std::array<float, 9> K;
    std::memcpy(K.data(), K9, sizeof(float) * 9);
    as<LandmarkEstimatorPoseBasedSmootherCUDA>(estimator, "LandmarkEstimatorPoseBasedSmoother")->setCameraMatrix(K);
    return 0;
  });
}

int psp_landmark_smoother_compute_batch(psp_module* estimator, int n, float* state_world, int* number_of_optimizations, int n_frames,
                                        const float* frames_sensor_in_world, const int* offsets, const int* hist_frame,
                                        const float* hist_uv, const float* hist_point_in_camera, float* coords_in_local_map,
                                        uint8_t* inlier) {
  return guard([&] {
    return as<LandmarkEstimatorPoseBasedSmootherCUDA>(estimator, "LandmarkEstimatorPoseBasedSmoother")
      ->computeBatch(n, state_world, number_of_optimizations, n_frames, frames_sensor_in_world, offsets, hist_frame, hist_uv,
                     hist_point_in_camera, coords_in_local_map, inlier);
  });
}

}  // extern "C"
