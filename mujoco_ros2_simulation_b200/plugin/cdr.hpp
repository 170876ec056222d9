// cdr.hpp -- the messages of the sensor path as they go on the wire.
//
// ROS 2 middlewares (rmw_fastrtps, rmw_cyclonedds) serialise messages as classic CDR ("XCDR1"):
// a 4-byte encapsulation header {0x00, 0x01 = little endian, 0x00, 0x00}, then the fields in
// declaration order, every primitive aligned to its own size relative to the start of the body,
// strings as uint32 length (including the terminating NUL) + bytes, sequences as uint32 count +
// elements, fixed arrays inline.  In a real plugin rclcpp / rmw does this; here the stand-in
// publisher can show -- and the tests can check against an independent parser -- that what the
// wrappers publish has the wire layout of sensor_msgs/LaserScan, Image, CameraInfo and
// PointCloud2 (ref:src/mujoco_lidar.cpp:275-279, ref:src/mujoco_cameras.cpp:253-264).
// Not used with -DB2_HAVE_ROS2 (the middleware owns serialisation there).
#ifndef B2_CDR_HPP
#define B2_CDR_HPP

#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "b2_compat.hpp"

namespace b2cdr
{

class Writer
{
public:
  Writer()
  {
    buf_ = { 0x00, 0x01, 0x00, 0x00 };  // CDR_LE, no options
  }
  template <class T>
  void prim(T v)
  {
    align(sizeof(T));
    const size_t at = buf_.size();
    buf_.resize(at + sizeof(T));
    std::memcpy(&buf_[at], &v, sizeof(T));
  }
  void str(const std::string& s)
  {
    prim<uint32_t>(static_cast<uint32_t>(s.size() + 1));
    buf_.insert(buf_.end(), s.begin(), s.end());
    buf_.push_back(0);
  }
  template <class T>
  void seq(const std::vector<T>& v)
  {
    prim<uint32_t>(static_cast<uint32_t>(v.size()));
    raw(v.data(), v.size());
  }
  template <class T>
  void raw(const T* p, size_t n)
  {
    if (n == 0)
    {
      return;
    }
    align(sizeof(T));
    const size_t at = buf_.size();
    buf_.resize(at + n * sizeof(T));
    std::memcpy(&buf_[at], p, n * sizeof(T));
  }
  std::vector<uint8_t> take()
  {
    return std::move(buf_);
  }

private:
  void align(size_t a)
  {
    const size_t body = buf_.size() - 4;  // alignment is relative to the end of the encapsulation header
    const size_t pad = (a - body % a) % a;
    buf_.insert(buf_.end(), pad, 0);
  }
  std::vector<uint8_t> buf_;
};

inline void put(Writer& w, const std_msgs::msg::Header& h)
{
  w.prim<int32_t>(h.stamp.sec);
  w.prim<uint32_t>(h.stamp.nanosec);
  w.str(h.frame_id);
}

inline std::vector<uint8_t> serialize(const sensor_msgs::msg::LaserScan& m)
{
  Writer w;
  put(w, m.header);
  for (float f : { m.angle_min, m.angle_max, m.angle_increment, m.time_increment, m.scan_time, m.range_min,
                   m.range_max })
  {
    w.prim<float>(f);
  }
  w.seq(m.ranges);
  w.seq(m.intensities);
  return w.take();
}

inline std::vector<uint8_t> serialize(const sensor_msgs::msg::Image& m)
{
  Writer w;
  put(w, m.header);
  w.prim<uint32_t>(m.height);
  w.prim<uint32_t>(m.width);
  w.str(m.encoding);
  w.prim<uint8_t>(m.is_bigendian);
  w.prim<uint32_t>(m.step);
  w.seq(m.data);
  return w.take();
}

inline std::vector<uint8_t> serialize(const sensor_msgs::msg::CameraInfo& m)
{
  Writer w;
  put(w, m.header);
  w.prim<uint32_t>(m.height);
  w.prim<uint32_t>(m.width);
  w.str(m.distortion_model);
  w.seq(m.d);
  w.raw(m.k.data(), m.k.size());
  w.raw(m.r.data(), m.r.size());
  w.raw(m.p.data(), m.p.size());
  w.prim<uint32_t>(0);  // binning_x
  w.prim<uint32_t>(0);  // binning_y
  for (int i = 0; i < 4; ++i)
  {
    w.prim<uint32_t>(0);  // roi: x_offset, y_offset, height, width
  }
  w.prim<uint8_t>(0);  // roi.do_rectify
  return w.take();
}

inline std::vector<uint8_t> serialize(const sensor_msgs::msg::PointCloud2& m)
{
  Writer w;
  put(w, m.header);
  w.prim<uint32_t>(m.height);
  w.prim<uint32_t>(m.width);
  w.prim<uint32_t>(static_cast<uint32_t>(m.fields.size()));
  for (const auto& f : m.fields)
  {
    w.str(f.name);
    w.prim<uint32_t>(f.offset);
    w.prim<uint8_t>(f.datatype);
    w.prim<uint32_t>(f.count);
  }
  w.prim<uint8_t>(m.is_bigendian ? 1 : 0);
  w.prim<uint32_t>(m.point_step);
  w.prim<uint32_t>(m.row_step);
  w.seq(m.data);
  w.prim<uint8_t>(m.is_dense ? 1 : 0);
  return w.take();
}

}  // namespace b2cdr
#endif
