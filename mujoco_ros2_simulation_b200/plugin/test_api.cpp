// test_api.cpp -- C entry points that let pytest (ctypes) drive the C++ wrapper classes the way
// MujocoSystemInterface does: construct + register in on_init, init/update in on_activate, close in
// the destructor (ref:src/mujoco_system_interface.cpp:707-717, 938-939, 467-476).  Only compiled
// with the stand-in ROS 2 / MuJoCo types of b2_compat.hpp.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstring>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "cdr.hpp"
#include "mujoco_cameras.hpp"
#include "mujoco_lidar.hpp"

using namespace mujoco_ros2_control;

namespace
{
struct Harness
{
  mjModel* m;
  mjData* d;
  std::recursive_mutex mtx;
  rclcpp::Node::SharedPtr node = std::make_shared<rclcpp::Node>();
  hardware_interface::HardwareInfo info;
  std::unique_ptr<SceneMirror> mirror;
  std::unique_ptr<MujocoLidar> lidar;
  std::unique_ptr<MujocoCameras> cameras;
  std::string error;
  double lidar_rate = 5.0, camera_rate = 5.0;  // the interface's defaults (ref :558-560)
};

// "sensor|key=val;key=val\nsensor2|..." -> HardwareInfo (what the URDF <ros2_control> block carries)
void parse_info(const char* text, hardware_interface::HardwareInfo& info)
{
  std::istringstream in(text ? text : "");
  std::string line;
  while (std::getline(in, line))
  {
    const auto bar = line.find('|');
    if (bar == std::string::npos)
    {
      continue;
    }
    hardware_interface::ComponentInfo c;
    c.name = line.substr(0, bar);
    std::istringstream kv(line.substr(bar + 1));
    std::string item;
    while (std::getline(kv, item, ';'))
    {
      const auto eq = item.find('=');
      if (eq != std::string::npos)
      {
        c.parameters[item.substr(0, eq)] = item.substr(eq + 1);
      }
    }
    info.sensors.push_back(c);
  }
}
}  // namespace

extern "C" {

// on_init part 1: wrappers constructed and registered (no GPU needed)
void* b2p_create(mjModel* m, mjData* d, const char* hardware_info, double lidar_rate, double camera_rate)
{
  auto* h = new Harness();
  h->m = m;
  h->d = d;
  parse_info(hardware_info, h->info);
  h->lidar_rate = lidar_rate;
  h->camera_rate = camera_rate;
  h->lidar = std::make_unique<MujocoLidar>(h->node, &h->mtx, d, m, lidar_rate, nullptr);
  h->cameras = std::make_unique<MujocoCameras>(h->node, &h->mtx, d, m, camera_rate, nullptr);
  return h;
}

int b2p_register(void* hp)
{
  auto* h = static_cast<Harness*>(hp);
  try
  {
    h->cameras->register_cameras(h->info);
  }
  catch (const std::exception& e)
  {
    h->error = std::string("register_cameras: ") + e.what();
    return -2;
  }
  if (!h->lidar->register_lidar(h->info))
  {
    h->error = h->lidar->last_error();
    return -1;  // on_init returns CallbackReturn::FAILURE here (ref :713-717)
  }
  return 0;
}

// on_init part 2 + on_activate: scene mirror + GPU objects (needs a B200)
int b2p_activate(void* hp)
{
  auto* h = static_cast<Harness*>(hp);
  try
  {
    h->mirror = std::make_unique<SceneMirror>(h->m, 0);
  }
  catch (const std::exception& e)
  {
    h->error = e.what();
    return -1;
  }
  // re-create the wrappers with the mirror, as the interface would construct them
  h->lidar = std::make_unique<MujocoLidar>(h->node, &h->mtx, h->d, h->m, h->lidar_rate, h->mirror.get());
  h->cameras = std::make_unique<MujocoCameras>(h->node, &h->mtx, h->d, h->m, h->camera_rate, h->mirror.get());
  h->cameras->register_cameras(h->info);
  if (!h->lidar->register_lidar(h->info))
  {
    h->error = h->lidar->last_error();
    return -1;
  }
  if (!h->lidar->create_engines())
  {
    h->error = h->lidar->last_error();
    return -2;
  }
  if (!h->cameras->create_engines())
  {
    h->error = h->cameras->last_error();
    return -3;
  }
  return 0;
}

// one publish cycle of both wrappers
int b2p_update(void* hp)
{
  auto* h = static_cast<Harness*>(hp);
  h->lidar->update();
  if (!h->lidar->last_error().empty())
  {
    h->error = h->lidar->last_error();
    return -1;
  }
  h->cameras->update();
  if (!h->cameras->last_error().empty())
  {
    h->error = h->cameras->last_error();
    return -2;
  }
  return 0;
}

// threads: init() starts them at the publish rate, close() joins them
int b2p_run_threads(void* hp, int milliseconds)
{
  auto* h = static_cast<Harness*>(hp);
  h->lidar->init();
  h->cameras->init();
  std::this_thread::sleep_for(std::chrono::milliseconds(milliseconds));
  h->lidar->close();
  h->cameras->close();
  return 0;
}

const char* b2p_error(void* hp)
{
  return static_cast<Harness*>(hp)->error.c_str();
}

int b2p_lidar_count(void* hp)
{
  return (int)static_cast<Harness*>(hp)->lidar->lidars().size();
}

// static LaserScan fields + bookkeeping of lidar k
int b2p_lidar_info(void* hp, int k, int* num_rangefinders, float* f7, int* bodyid, char* frame, char* topic,
                   int* sensor_indexes, double* local_dir)
{
  const auto& l = static_cast<Harness*>(hp)->lidar->lidars().at(k);
  *num_rangefinders = l.num_rangefinders;
  const auto& msg = l.laser_scan_msg;
  f7[0] = msg.angle_min;
  f7[1] = msg.angle_max;
  f7[2] = msg.angle_increment;
  f7[3] = msg.time_increment;
  f7[4] = msg.scan_time;
  f7[5] = msg.range_min;
  f7[6] = msg.range_max;
  *bodyid = l.bodyid;
  std::strncpy(frame, msg.header.frame_id.c_str(), 127);
  std::strncpy(topic, l.scan_pub ? l.scan_pub->get_topic_name().c_str() : "", 127);
  for (int i = 0; i < l.num_rangefinders; ++i)
  {
    sensor_indexes[i] = l.sensor_indexes[i];
  }
  std::memcpy(local_dir, l.local_dir.data(), l.local_dir.size() * sizeof(double));
  return 0;
}

// last published LaserScan of lidar k
int b2p_lidar_scan(void* hp, int k, float* ranges, int* n_published, int* n_intensities)
{
  const auto& l = static_cast<Harness*>(hp)->lidar->lidars().at(k);
  const auto& msg = l.scan_pub->last();
  std::memcpy(ranges, msg.ranges.data(), msg.ranges.size() * sizeof(float));
  *n_published = (int)l.scan_pub->count();
  *n_intensities = (int)msg.intensities.size();
  return (int)msg.ranges.size();
}

int b2p_camera_count(void* hp)
{
  return (int)static_cast<Harness*>(hp)->cameras->cameras().size();
}

int b2p_camera_info(void* hp, int k, int* wh, double* K, double* R, double* P, int* nd, char* frame,
                    char* topics /* 4 x 128 */, char* encodings /* 2 x 16 */, int* steps)
{
  auto& c = static_cast<Harness*>(hp)->cameras->cameras().at(k);
  wh[0] = (int)c.width;
  wh[1] = (int)c.height;
  std::memcpy(K, c.camera_info.k.data(), 9 * sizeof(double));
  std::memcpy(R, c.camera_info.r.data(), 9 * sizeof(double));
  std::memcpy(P, c.camera_info.p.data(), 12 * sizeof(double));
  *nd = (int)c.camera_info.d.size();
  std::strncpy(frame, c.frame_name.c_str(), 127);
  std::strncpy(topics, c.info_topic.c_str(), 127);
  std::strncpy(topics + 128, c.image_topic.c_str(), 127);
  std::strncpy(topics + 256, c.depth_topic.c_str(), 127);
  std::strncpy(topics + 384, c.points_topic.c_str(), 127);
  std::strncpy(encodings, c.image.encoding.c_str(), 15);
  std::strncpy(encodings + 16, c.depth_image.encoding.c_str(), 15);
  steps[0] = (int)c.image.step;
  steps[1] = (int)c.depth_image.step;
  steps[2] = (int)c.cloud.point_step;
  steps[3] = (int)c.cloud.row_step;
  return 0;
}

// last published depth image / cloud / raw GL depth of camera k
int b2p_camera_frames(void* hp, int k, float* gl_depth, float* depth, float* cloud)
{
  auto& c = static_cast<Harness*>(hp)->cameras->cameras().at(k);
  const size_t n = (size_t)c.width * c.height;
  if (gl_depth)
  {
    std::memcpy(gl_depth, c.depth_buffer.data(), n * sizeof(float));
  }
  if (depth)
  {
    std::memcpy(depth, c.depth_image_pub->last().data.data(), n * sizeof(float));
  }
  if (cloud)
  {
    std::memcpy(cloud, c.cloud_pub->last().data.data(), n * 12);
  }
  return (int)c.depth_image_pub->count();
}

// feed a GL depth buffer instead of ray casting it (what mjr_readPixels would deliver)
int b2p_camera_set_gl_depth(void* hp, const float* gl_depth)
{
  auto* h = static_cast<Harness*>(hp);
  h->cameras->set_gl_readback([gl_depth](CameraData& cam) {
    std::memcpy(cam.depth_buffer.data(), gl_depth, cam.depth_buffer.size() * sizeof(float));
    for (size_t i = 0; i < cam.image_buffer.size(); ++i)
    {
      cam.image_buffer[i] = (uint8_t)(i * 7);  // a recognisable RGB pattern for the flip check
    }
    return true;
  });
  return 0;
}

int b2p_camera_rgb(void* hp, int k, uint8_t* rgb_published, uint8_t* rgb_gl)
{
  auto& c = static_cast<Harness*>(hp)->cameras->cameras().at(k);
  std::memcpy(rgb_published, c.image_pub->last().data.data(), c.image.data.size());
  std::memcpy(rgb_gl, c.image_buffer.data(), c.image_buffer.size());
  return (int)c.image.data.size();
}

// A9 / SURVEY B3: what the physics thread pays for SceneMirror::refresh while a sensor thread keeps
// the GPU busy.  Thread A casts an (n_el x n_az)-ray lidar back to back through the blocking ABI;
// thread B plays PhysicsLoop at `hz`: lock the sim mutex, move the geoms a little ("mj_step"),
// refresh the mirror, unlock.  out = {p50, p99, max refresh time in microseconds, casts completed,
// refreshes done}.  Returns 0, or a negative code with b2p_error set.
int b2p_refresh_stress(void* hp, int n_el, int n_az, double hz, double seconds, double* out)
{
  auto* h = static_cast<Harness*>(hp);
  if (!h->mirror)
  {
    h->error = "b2p_refresh_stress before b2p_activate";
    return -1;
  }
  // a 3-D fan around the first lidar's body frame
  std::vector<double> dirs((size_t)n_el * n_az * 3);
  for (int e = 0; e < n_el; ++e)
  {
    const double el = -0.4 + 0.6 * e / std::max(1, n_el - 1);
    for (int a = 0; a < n_az; ++a)
    {
      const double az = 6.283185307179586 * a / n_az;
      double* d = &dirs[((size_t)e * n_az + a) * 3];
      d[0] = std::cos(el) * std::cos(az);
      d[1] = std::cos(el) * std::sin(az);
      d[2] = std::sin(el);
    }
  }
  b2_lidar_desc desc{};
  desc.nray = n_el * n_az;
  desc.local_dir = dirs.data();
  desc.tile_w = n_az;
  desc.flg_static = 1;
  desc.bodyexclude = -1;
  desc.range_max = 1000.0;
  b2_lidar* big = nullptr;
  {
    std::unique_lock<std::recursive_mutex> lock(h->mtx);
    h->mirror->refresh(h->d);  // first poses (topology) if nobody did it yet
  }
  if (b2_lidar_create(h->mirror->handle(), &desc, &big) != B2_OK)
  {
    h->error = b2_last_error();
    return -2;
  }
  std::atomic_bool stop{ false };
  std::atomic_int casts{ 0 }, cast_fail{ 0 };
  std::thread sensor([&]() {
    std::vector<float> ranges((size_t)desc.nray);
    const double pos[3] = { 0.5, 0.5, 0.8 };
    const double mat[9] = { 1, 0, 0, 0, 1, 0, 0, 0, 1 };
    while (!stop)
    {
      if (b2_lidar_cast(big, 1, nullptr, pos, mat, ranges.data(), nullptr, nullptr, nullptr) != B2_OK)
      {
        cast_fail++;
        break;
      }
      casts++;
    }
  });
  std::vector<double> us;
  const auto t_start = std::chrono::steady_clock::now();
  const auto period = std::chrono::duration<double>(1.0 / hz);
  int tick = 0;
  while (std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count() < seconds)
  {
    {
      std::unique_lock<std::recursive_mutex> lock(h->mtx);
      for (int g = 0; g < h->m->ngeom; ++g)
      {
        h->d->geom_xpos[3 * g] += (tick & 1) ? 1e-4 : -1e-4;  // the stand-in for mj_step
      }
      const auto a = std::chrono::steady_clock::now();
      h->mirror->refresh(h->d);
      const auto b = std::chrono::steady_clock::now();
      us.push_back(std::chrono::duration<double, std::micro>(b - a).count());
    }
    ++tick;
    std::this_thread::sleep_until(t_start + tick * period);
  }
  stop = true;
  sensor.join();
  b2_lidar_destroy(big);
  if (cast_fail)
  {
    h->error = std::string("cast failed: ") + b2_last_error();
    return -3;
  }
  std::sort(us.begin(), us.end());
  out[0] = us.empty() ? 0 : us[us.size() / 2];
  out[1] = us.empty() ? 0 : us[(size_t)((us.size() - 1) * 0.99)];
  out[2] = us.empty() ? 0 : us.back();
  out[3] = casts;
  out[4] = (double)us.size();
  return 0;
}

// what the wrappers logged so far (stand-in rclcpp logger), newline separated; returns the line count
int b2p_logs(void* hp, char* out, int cap)
{
  auto* h = static_cast<Harness*>(hp);
  const auto& lines = *h->node->get_logger().lines;
  std::string all;
  for (const auto& l : lines)
  {
    all += l;
    all += '\n';
  }
  if (out != nullptr && cap > 0)
  {
    std::strncpy(out, all.c_str(), (size_t)cap - 1);
    out[cap - 1] = 0;
  }
  return (int)lines.size();
}

// the last published message of a publisher as it would go on the wire (classic CDR, little endian):
// which = 0 LaserScan of lidar k, 1 depth Image, 2 CameraInfo, 3 PointCloud2, 4 colour Image of camera k.
// Returns the byte count (call with out = nullptr to size the buffer).
long b2p_wire_bytes(void* hp, int which, int k, uint8_t* out, long cap)
{
  auto* h = static_cast<Harness*>(hp);
  std::vector<uint8_t> bytes;
  if (which == 0)
  {
    bytes = b2cdr::serialize(h->lidar->lidars().at(k).scan_pub->last());
  }
  else
  {
    auto& c = h->cameras->cameras().at(k);
    bytes = which == 1 ? b2cdr::serialize(c.depth_image_pub->last()) :
            which == 2 ? b2cdr::serialize(c.camera_info_pub->last()) :
            which == 3 ? b2cdr::serialize(c.cloud_pub->last()) :
                         b2cdr::serialize(c.image_pub->last());
  }
  if (out != nullptr && cap >= (long)bytes.size())
  {
    std::memcpy(out, bytes.data(), bytes.size());
  }
  return (long)bytes.size();
}

// publish counts of camera k: image, depth, camera_info, points
int b2p_camera_pub_counts(void* hp, int k, int* out4)
{
  auto& c = static_cast<Harness*>(hp)->cameras->cameras().at(k);
  out4[0] = (int)c.image_pub->count();
  out4[1] = (int)c.depth_image_pub->count();
  out4[2] = (int)c.camera_info_pub->count();
  out4[3] = (int)c.cloud_pub->count();
  return 0;
}

void b2p_destroy(void* hp)
{
  auto* h = static_cast<Harness*>(hp);
  h->lidar.reset();
  h->cameras.reset();
  h->mirror.reset();
  delete h;
}

size_t b2p_sizeof_mjmodel()
{
  return sizeof(mjModel);
}
size_t b2p_sizeof_mjdata()
{
  return sizeof(mjData);
}

}  // extern "C"
