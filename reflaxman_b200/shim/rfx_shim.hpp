// rfx_shim.hpp — host-side C++ mirror of the reference's Render / Scene / Camera class surface, implemented on top of
// the C ABI (include/reflax_c.h).  With the forwarding headers next to this file (Render.h, Scene.h, Camera.h, ...)
// the reference's own Pulse.cpp and front ends (src/linux/main.cpp, src/windows/Main.cpp) compile UNCHANGED against
// the B200 path: same class names, same public members, same argument meaning, same error behaviour (bool returns
// and silent fall-backs, never exceptions).  Nothing here traces a ray on the CPU — there is no CPU fallback.
//
// What each class stands in for (reference path:line under src/common):
//   Vector3, Matrix33, Color, Material   value types            Vector3.h, Matrix33.h, Color.h, Material.h
//   Texture                              ARGB image + TGA/BMP   Texture.h:6-34, Texture.cpp:34-214, image_headers.h
//   Sphere, Triangle, OmniLight          handles returned by Scene::add*   Scene.h:33-35, Triangle.h:26-27
//   Scene                                scene description      Scene.h:12-40, Scene.cpp:10-71
//   Camera                               eye/view/fov + UI kinematics      Camera.h:31-63, Camera.cpp:24-56,110-248
//   Render                               the drop-in boundary   Render.h:7-42, Render.cpp:5-226
//
// Design: Scene is a plain host-side description (it has to be assignable — `scene = Scene(colour, power)` in
// Render::loadScene — and triangles are mutated after insertion through the pointers addTriangle returns), with a
// revision counter; Render re-issues it through rfx_scene_reset/rfx_add_* whenever the revision it uploaded is stale.
#pragma once

#include <assert.h>
#include <float.h>
#include <math.h>
#include <memory>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

#include "reflax_c.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#ifndef M_PI_2
#define M_PI_2 1.57079632679489661923
#endif

// ---------------------------------------------------------------------------------------------------- trace_math.h
namespace Tracemath
{
  const float VERY_SMALL_NUMBER = 1.08420217248550443e-19f;   // sqrtf(FLT_MIN)
  const float DELTA = 0.0001f;
  template <typename T> inline T clamp(T v, T lo, T hi) { return v < lo ? lo : (v > hi ? hi : v); }
  inline int min(int a, int b) { return a < b ? a : b; }
  inline int max(int a, int b) { return a > b ? a : b; }
  inline int min(unsigned a, unsigned b) { return int(a < b ? a : b); }   // the reference returns int here too
  inline int max(unsigned a, unsigned b) { return int(a > b ? a : b); }
  inline float min(float a, float b) { return a < b ? a : b; }
  inline float max(float a, float b) { return a > b ? a : b; }
}
#ifndef FORBIDE_USING_AIRLY_NAMESPACE
using namespace Tracemath;
#endif

// ------------------------------------------------------------------------------------------------------- Vector3.h
class Matrix33;
class Vector3
{
public:
  float x, y, z;
  Vector3() {}
  Vector3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
  float sqLength() const { return x * x + y * y + z * z; }
  float length() const { return sqrtf(sqLength()); }
  Vector3 operator+(const Vector3 & o) const { return Vector3(x + o.x, y + o.y, z + o.z); }
  Vector3 operator-(const Vector3 & o) const { return Vector3(x - o.x, y - o.y, z - o.z); }
  Vector3 operator-() const { return Vector3(-x, -y, -z); }
  Vector3 operator*(float f) const { return Vector3(x * f, y * f, z * f); }
  friend Vector3 operator*(float f, const Vector3 & v) { return Vector3(v.x * f, v.y * f, v.z * f); }
  float operator*(const Vector3 & o) const { return x * o.x + y * o.y + z * o.z; }                       // dot
  Vector3 operator%(const Vector3 & o) const { return Vector3(y * o.z - z * o.y, z * o.x - x * o.z, x * o.y - y * o.x); }   // cross
  Vector3 operator/(float f) const { return fabsf(f) > Tracemath::VERY_SMALL_NUMBER ? Vector3(x / f, y / f, z / f) : *this; }
  Vector3 & operator+=(const Vector3 & o) { x += o.x; y += o.y; z += o.z; return *this; }
  Vector3 & operator-=(const Vector3 & o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
  Vector3 & operator*=(float f) { x *= f; y *= f; z *= f; return *this; }
  Vector3 & operator/=(float f) { if (fabsf(f) > Tracemath::VERY_SMALL_NUMBER) { x /= f; y /= f; z /= f; } return *this; }
  bool operator==(const Vector3 & o) const { return x == o.x && y == o.y && z == o.z; }
  bool operator!=(const Vector3 & o) const { return !(*this == o); }
  Vector3 normalized() const { const float l = length(); return l > Tracemath::VERY_SMALL_NUMBER ? *this / l : *this; }
  void normalize() { *this = normalized(); }
};
namespace Tracemath
{
  inline Vector3 normalize(const Vector3 & v) { return v.normalized(); }
}

// ------------------------------------------------------------------------------------------------------ Matrix33.h
class Matrix33
{
public:
  float _11, _12, _13, _21, _22, _23, _31, _32, _33;   // row-major, as the reference's anonymous struct
  Matrix33() {}
  Matrix33(const Vector3 & u, const Vector3 & v, const Vector3 & n)   // columns
    : _11(u.x), _12(v.x), _13(n.x), _21(u.y), _22(v.y), _23(n.y), _31(u.z), _32(v.z), _33(n.z) {}
  Matrix33(float a, float b, float c, float d, float e, float f, float g, float h, float i)
    : _11(a), _12(b), _13(c), _21(d), _22(e), _23(f), _31(g), _32(h), _33(i) {}
  Vector3 getCol(int i) const { const float * m = &_11; return Vector3(m[i], m[3 + i], m[6 + i]); }
  void setCol(int i, const Vector3 & v) { float * m = &_11; m[i] = v.x; m[3 + i] = v.y; m[6 + i] = v.z; }
  Vector3 operator*(const Vector3 & v) const
  {
    return Vector3(v.x * _11 + v.y * _12 + v.z * _13, v.x * _21 + v.y * _22 + v.z * _23, v.x * _31 + v.y * _32 + v.z * _33);
  }
  static Matrix33 makeRotation(float yaw, float pitch)   // reference Matrix33.cpp:270-282
  {
    const float ys = sin(yaw), yc = cos(yaw), ps = sin(pitch), pc = cos(pitch);
    const Vector3 front(ys * pc, ps, yc * pc);
    const Vector3 right = Vector3(0.0f, 1.0f, 0.0f) % front;
    const Vector3 up = front % right;
    return Matrix33(right.normalized(), up.normalized(), front.normalized());
  }
};

// --------------------------------------------------------------------------------------------------------- Color.h
typedef uint32_t ARGB;
#define MAKEARGB(a, r, g, b) ((ARGB)((((unsigned char)(a)) & 0xffu) << 24 | (((unsigned char)(r)) & 0xffu) << 16 | \
                                     (((unsigned char)(g)) & 0xffu) << 8 | (((unsigned char)(b)) & 0xffu)))
#define ARGB_RED(c) (((c) >> 16) & 0xFF)
#define ARGB_GREEN(c) (((c) >> 8) & 0xFF)
#define ARGB_BLUE(c) ((c) & 0xFF)
#define ARGB_ALPHA(c) (((c) >> 24) & 0xFF)
class Color
{
public:
  float r, g, b;
  Color() {}
  Color(float r_, float g_, float b_) : r(r_), g(g_), b(b_) {}
  Color(const ARGB & c) : r(float(ARGB_RED(c)) / 255.0f), g(float(ARGB_GREEN(c)) / 255.0f), b(float(ARGB_BLUE(c)) / 255.0f) {}
  Color operator*(const float & f) const { return Color(r * f, g * f, b * f); }
  friend Color operator*(const float & f, const Color & c) { return Color(c.r * f, c.g * f, c.b * f); }
  Color operator*(const Color & o) const { return Color(r * o.r, g * o.g, b * o.b); }
  Color operator+(const Color & o) const { return Color(r + o.r, g + o.g, b + o.b); }
  Color operator-(const Color & o) const { return Color(r - o.r, g - o.g, b - o.b); }
  Color operator/(const float & f) const { return fabsf(f) > Tracemath::VERY_SMALL_NUMBER ? Color(r / f, g / f, b / f) : *this; }
  Color & operator+=(const Color & o) { r += o.r; g += o.g; b += o.b; return *this; }
  Color & operator*=(const Color & o) { r *= o.r; g *= o.g; b *= o.b; return *this; }
  Color & operator*=(const float & f) { r *= f; g *= f; b *= f; return *this; }
  Color & operator/=(const float & f) { if (fabsf(f) > Tracemath::VERY_SMALL_NUMBER) { r /= f; g /= f; b /= f; } return *this; }
  ARGB argb() const { return MAKEARGB(0, r * 255.999f, g * 255.999f, b * 255.999f); }   // reference Color.cpp:114-117
  void clamp() { r = Tracemath::clamp(r, 0.0f, 1.0f); g = Tracemath::clamp(g, 0.0f, 1.0f); b = Tracemath::clamp(b, 0.0f, 1.0f); }
};

// ------------------------------------------------------------------------------------------------------ Material.h
class Material
{
public:
  enum Type { mtMetal, mtDielectric };
  Type type;
  Color color;
  float reflectivity, transparency;
  Material() {}
  Material(Type t, const Color & c, float refl, float transp)
    : type(t), color(c), reflectivity(Tracemath::clamp(refl, 0.0f, 1.0f)), transparency(Tracemath::clamp(transp, 0.0f, 1.0f)) {}
};

// ------------------------------------------------------------------------------------------------------- Texture.h
// File formats (reference Texture.cpp:34-173, image_headers.h): type-2 uncompressed TGA, 24/32 bpp, rows stored in file
// order (the origin bit is ignored); 32-bpp bottom-up BMP and 32-bpp TGA on save.
class Texture
{
  unsigned int width, height;
  std::vector<ARGB> colorBuf;
  static void put16(unsigned char * p, unsigned v) { p[0] = (unsigned char)(v & 0xff); p[1] = (unsigned char)((v >> 8) & 0xff); }
  static void put32(unsigned char * p, uint32_t v) { put16(p, v & 0xffff); put16(p + 2, v >> 16); }
public:
  Texture() : width(0), height(0) {}
  Texture(unsigned int w, unsigned int h) { resize(w, h); }
  Texture(const char * fileName) : width(0), height(0) { loadFromFile(fileName); }

  static bool hasExt(const char * fileName, const char * ext)
  {
    const char * dot = strrchr(fileName, '.');
    return dot && !strcmp(dot, ext);
  }
  bool loadFromFile(const char * fileName) { return hasExt(fileName, ".tga") ? loadFromTGAFile(fileName) : false; }
  bool saveToFile(const char * fileName) const
  {
    if (hasExt(fileName, ".tga")) return saveToTGAFile(fileName);
    if (hasExt(fileName, ".bmp")) return saveToBMPFile(fileName);
    return false;
  }
  bool loadFromTGAFile(const char * fileName)
  {
    bool ok = false;
    if (FILE * f = fopen(fileName, "rb"))
    {
      unsigned char h[18];
      if (fread(h, sizeof(h), 1, f) == 1 && h[2] == 2)
      {
        const unsigned w = h[12] | (h[13] << 8), ht = h[14] | (h[15] << 8), bpp = h[16];
        const unsigned cmlen = h[5] | (h[6] << 8), cmbits = h[7];
        const long dataAt = long(sizeof(h)) + h[0] + long(cmlen) * cmbits / 8;
        if ((bpp == 24 || bpp == 32) && !fseek(f, dataAt, SEEK_SET))
        {
          const size_t n = size_t(w) * ht, px = bpp / 8;
          std::vector<unsigned char> raw(n * px);
          if (n && fread(raw.data(), px, n, f) == n)
          {
            colorBuf.resize(n);
            for (size_t i = 0; i < n; i++)
            {
              const unsigned char * p = &raw[i * px];
              colorBuf[i] = MAKEARGB(px == 4 ? p[3] : 0xFF, p[2], p[1], p[0]);
            }
            width = w; height = ht;
            ok = true;
          }
        }
      }
      fclose(f);
    }
    if (!ok) { colorBuf.clear(); width = height = 0; }
    return ok;
  }
  bool saveToTGAFile(const char * fileName) const
  {
    FILE * f = fopen(fileName, "wb");
    if (!f) return false;
    unsigned char h[18] = { 0 };
    h[2] = 2; put16(h + 12, width); put16(h + 14, height); h[16] = 32;
    const bool ok = fwrite(h, sizeof(h), 1, f) == 1 && fwrite(colorBuf.data(), size_t(width) * height * 4, 1, f) == 1;
    fclose(f);
    return ok;
  }
  bool saveToBMPFile(const char * fileName) const
  {
    FILE * f = fopen(fileName, "wb");
    if (!f) return false;
    unsigned char h[54] = { 0 };
    h[0] = 'B'; h[1] = 'M';
    put32(h + 2, 54 + width * height * 4); put32(h + 10, 54);
    put32(h + 14, 40); put32(h + 18, width); put32(h + 22, height); put16(h + 26, 1); put16(h + 28, 32);
    const bool ok = fwrite(h, sizeof(h), 1, f) == 1 && fwrite(colorBuf.data(), size_t(width) * height * 4, 1, f) == 1;
    fclose(f);
    return ok;
  }
  ARGB * getColorBuffer() const { return const_cast<ARGB *>(colorBuf.data()); }
  void resize(unsigned int w, unsigned int h) { width = w; height = h; colorBuf.resize(size_t(w) * h); }
  void clear(ARGB c) { for (size_t i = 0; i < colorBuf.size(); i++) colorBuf[i] = c; }
  unsigned int getWidth() const { return width; }
  unsigned int getHeight() const { return height; }
  bool empty() const { return colorBuf.empty(); }
};

// ------------------------------------------------------------------------------------- Scene objects (handles) / Scene
class Scene;
class Render;
class SceneObject { public: virtual ~SceneObject() {} };
class Sphere : public SceneObject
{
public:
  Vector3 center; float radius; Material material;
};
class Triangle : public SceneObject
{
  friend class Scene; friend class Render;
  Scene * owner; Vector3 v[3]; Material material; const Texture * texture; float uv[6];
public:
  Triangle() : owner(NULL), texture(NULL) { memset(uv, 0, sizeof(uv)); }
  // reference Triangle.cpp:110-120 — called on the pointer addTriangle returned, AFTER insertion
  inline void setTexture(const Texture * tex, float u1, float v1, float u2, float v2, float u3, float v3);
};
class OmniLight
{
public:
  Vector3 origin; float radius; Color color; float power;
};

class Scene
{
  friend class Render; friend class Triangle;
  struct Data
  {
    Color ambient; float ambientPower;
    std::vector<std::unique_ptr<SceneObject> > objects;   // insertion order (closest-hit ties, reference Scene.cpp:98)
    std::vector<std::unique_ptr<OmniLight> > lights;
    std::vector<std::unique_ptr<Texture> > textures;
    Texture skybox;
    unsigned revision;
    Data() : ambient(0, 0, 0), ambientPower(0), revision(1) {}
  };
  std::shared_ptr<Data> d;   // shallow-copy semantics like the reference's pointer vectors, without the double free
  Render * host;             // the Render this scene is a member of (so that Scene::trace can reach the device)
  void touch() { d->revision++; }
public:
  Scene() : d(new Data()), host(NULL) {}
  Scene(const Color & diffLightColor, float diffLightPower) : d(new Data()), host(NULL) { d->ambient = diffLightColor; d->ambientPower = diffLightPower; }
  Scene(const Scene & o) : d(o.d), host(NULL) {}
  Scene & operator=(const Scene & o) { d = o.d; return *this; }   // `scene = Scene(colour, power)` keeps its Render binding

  Sphere * addSphere(const Vector3 & center, float radius, const Material & material)
  {
    Sphere * s = new Sphere(); s->center = center; s->radius = radius; s->material = material;
    d->objects.push_back(std::unique_ptr<SceneObject>(s)); touch();
    return s;
  }
  Triangle * addTriangle(const Vector3 & v1, const Vector3 & v2, const Vector3 & v3, const Material & material)
  {
    Triangle * t = new Triangle(); t->owner = this; t->v[0] = v1; t->v[1] = v2; t->v[2] = v3; t->material = material;
    d->objects.push_back(std::unique_ptr<SceneObject>(t)); touch();
    return t;
  }
  OmniLight * addLight(const Vector3 & origin, float radius, const Color & color, float power)
  {
    OmniLight * l = new OmniLight(); l->origin = origin; l->radius = radius; l->color = color; l->power = power;
    d->lights.push_back(std::unique_ptr<OmniLight>(l)); touch();
    return l;
  }
  Texture * addTexture(const char * fileName)
  {
    d->textures.push_back(std::unique_ptr<Texture>(new Texture(fileName))); touch();
    return d->textures.back().get();
  }
  bool setSkyboxTexture(const char * fileName) { const bool ok = d->skybox.loadFromFile(fileName); touch(); return ok; }

  // Scene::trace (reference Scene.cpp:73): one ray through the device path; consumes one randDir from the stream
  inline Color trace(Vector3 origin, Vector3 ray, int reflNumber) const;
};

inline void Triangle::setTexture(const Texture * tex, float u1, float v1, float u2, float v2, float u3, float v3)
{
  texture = tex;
  uv[0] = u1; uv[1] = v1; uv[2] = u2; uv[3] = v2; uv[4] = u3; uv[5] = v3;
  if (owner) owner->touch();
}

// -------------------------------------------------------------------------------------------------------- Camera.h
enum Control
{
  turnLeftMask = 1 << 0, turnRightMask = 1 << 1, turnUpMask = 1 << 2, turnDownMask = 1 << 3,
  shiftLeftMask = 1 << 6, shiftRightMask = 1 << 7, shiftUpMask = 1 << 8, shiftDownMask = 1 << 9,
  shiftForwardMask = 1 << 10, shiftBackMask = 1 << 11,
};
namespace Default
{
  const float turnAccel = 2.0f, turnDecel = 2.0f, maxTurnSpeed = 0.2f;
  const float shiftAccel = 50.0f, shiftDecel = 50.0f, maxShiftSpeed = 10.0f;
}
class Camera
{
  // one axis of the keyboard kinematics (reference Camera.cpp:118-203): accelerate while a key is held (braking harder
  // when reversing a strafe), otherwise decelerate towards rest
  static float axis(float v, int dir, float accel, float decel, float vmax, float dt, bool brakeOnReverse)
  {
    if (dir > 0) return Tracemath::clamp(v + ((brakeOnReverse && v < 0.0f) ? decel + accel : accel) * dt, -vmax, vmax);
    if (dir < 0) return Tracemath::clamp(v - ((brakeOnReverse && v > 0.0f) ? decel + accel : accel) * dt, -vmax, vmax);
    if (v < 0.0f) return Tracemath::min(0.0f, v + decel * dt);
    if (v > 0.0f) return Tracemath::max(0.0f, v - decel * dt);
    return v;
  }
  static int pick(int flags, int plus, int minus) { const int f = flags & (plus | minus); return f == plus ? 1 : (f == minus ? -1 : 0); }
public:
  float turnRLSpeed, turnUDSpeed, shiftRLSpeed, shiftUDSpeed, shiftFBSpeed;
  float yaw, pitch;
  float fov;
  Vector3 eye;
  Matrix33 view;

  Camera() : turnRLSpeed(0), turnUDSpeed(0), shiftRLSpeed(0), shiftUDSpeed(0), shiftFBSpeed(0), yaw(0), pitch(0), fov(0) {}
  Camera(const Vector3 & eye_, const Vector3 & at, float fov_)   // reference Camera.cpp:24-42
    : turnRLSpeed(0), turnUDSpeed(0), shiftRLSpeed(0), shiftUDSpeed(0), shiftFBSpeed(0), fov(fov_), eye(eye_)
  {
    const Vector3 up(0.0f, 1.0f, 0.0f);
    const Vector3 oz = Tracemath::normalize(at - eye_);
    const Vector3 ox = Tracemath::normalize(up % oz);
    const Vector3 oy = Tracemath::normalize(oz % ox);
    view.setCol(0, ox); view.setCol(1, oy); view.setCol(2, oz);
    yaw = acos(ox.z);
    if (ox.x < 0) yaw = 2.0f * float(M_PI) - yaw;
    yaw -= M_PI_2;
    pitch = asin(oz.y);
  }
  // the reference's copy operations reset the speeds (Camera.cpp:58-106)
  Camera(const Camera & c) : turnRLSpeed(0), turnUDSpeed(0), shiftRLSpeed(0), shiftUDSpeed(0), shiftFBSpeed(0),
                             yaw(c.yaw), pitch(c.pitch), fov(c.fov), eye(c.eye), view(c.view) {}
  Camera & operator=(const Camera & c)
  {
    eye = c.eye; fov = c.fov; view = c.view; yaw = c.yaw; pitch = c.pitch;
    turnRLSpeed = turnUDSpeed = shiftRLSpeed = shiftUDSpeed = shiftFBSpeed = 0;
    return *this;
  }

  void proceedControl(int flags, float dt)
  {
    const float pTurnRL = turnRLSpeed, pTurnUD = turnUDSpeed, pRL = shiftRLSpeed, pUD = shiftUDSpeed, pFB = shiftFBSpeed;
    turnRLSpeed = axis(turnRLSpeed, pick(flags, turnRightMask, turnLeftMask), Default::turnAccel, Default::turnDecel, Default::maxTurnSpeed, dt, false);
    turnUDSpeed = axis(turnUDSpeed, pick(flags, turnUpMask, turnDownMask), Default::turnAccel, Default::turnDecel, Default::maxTurnSpeed, dt, false);
    shiftRLSpeed = axis(shiftRLSpeed, pick(flags, shiftRightMask, shiftLeftMask), Default::shiftAccel, Default::shiftDecel, Default::maxShiftSpeed, dt, true);
    shiftUDSpeed = axis(shiftUDSpeed, pick(flags, shiftUpMask, shiftDownMask), Default::shiftAccel, Default::shiftDecel, Default::maxShiftSpeed, dt, false);
    shiftFBSpeed = axis(shiftFBSpeed, pick(flags, shiftForwardMask, shiftBackMask), Default::shiftAccel, Default::shiftDecel, Default::maxShiftSpeed, dt, false);

    const float twoPi = 2 * M_PI;
    yaw += dt * twoPi * (turnRLSpeed + pTurnRL) / 2.0f;                       // trapezoidal integration of the turn rates
    pitch = Tracemath::clamp(pitch + dt * twoPi * (turnUDSpeed + pTurnUD) / 2.0f, float(-0.95f * M_PI_2), float(0.95f * M_PI_2));
    if (yaw >= twoPi) yaw -= twoPi; else if (yaw <= -twoPi) yaw += twoPi;
    view = Matrix33::makeRotation(yaw, pitch);

    if (fabs(shiftRLSpeed) > FLT_EPSILON || fabs(shiftUDSpeed) > FLT_EPSILON || fabs(shiftFBSpeed) > FLT_EPSILON)
    {
      const Vector3 right = view.getCol(0), up(0.0f, 1.0f, 0.0f);
      const Vector3 front = Tracemath::normalize(right % up);
      Vector3 shift = 0.5f * (shiftRLSpeed + pRL) * right + 0.5f * (shiftUDSpeed + pUD) * up + 0.5f * (shiftFBSpeed + pFB) * front;
      const float sq = shift.sqLength();
      if (sq > Default::maxShiftSpeed * Default::maxShiftSpeed) shift = shift * Default::maxShiftSpeed / sqrtf(sq);
      eye += shift * dt;
    }
  }
  bool inMotion() const
  {
    return fabs(turnRLSpeed) > FLT_EPSILON || fabs(turnUDSpeed) > FLT_EPSILON || fabs(shiftRLSpeed) > FLT_EPSILON ||
           fabs(shiftUDSpeed) > FLT_EPSILON || fabs(shiftFBSpeed) > FLT_EPSILON;
  }
};

// -------------------------------------------------------------------------------------------------------- Render.h
class Render
{
  rfx_ctx * ctx;
  unsigned uploadedRevision;
  const void * uploadedScene;
  mutable std::vector<float> cache;     // imagePixel() is called once per pixel per repaint (Pulse.cpp:455-458): bulk-read once
  mutable bool cacheValid;

  static int deviceFromEnv() { const char * s = getenv("RFX_DEVICE"); return s ? atoi(s) : 0; }

  bool syncScene()   // re-issue the scene description through the C ABI when it changed since the last upload
  {
    Scene::Data & sd = *scene.d;
    if (uploadedScene == &sd && uploadedRevision == sd.revision) return true;
    const float amb[3] = { sd.ambient.r, sd.ambient.g, sd.ambient.b };
    if (rfx_scene_reset(ctx, amb, sd.ambientPower) < 0) return false;
    std::vector<const Texture *> texPtr;
    for (size_t i = 0; i < sd.textures.size(); i++)
    {
      const Texture & t = *sd.textures[i];
      rfx_add_texture_argb(ctx, t.getWidth(), t.getHeight(), t.empty() ? NULL : t.getColorBuffer());
      texPtr.push_back(&t);
    }
    if (!sd.skybox.empty())
      rfx_set_skybox(ctx, rfx_add_texture_argb(ctx, sd.skybox.getWidth(), sd.skybox.getHeight(), sd.skybox.getColorBuffer()));
    for (size_t i = 0; i < sd.lights.size(); i++)
    {
      const OmniLight & l = *sd.lights[i];
      const float o[3] = { l.origin.x, l.origin.y, l.origin.z }, c[3] = { l.color.r, l.color.g, l.color.b };
      rfx_add_light(ctx, o, l.radius, c, l.power);
    }
    for (size_t i = 0; i < sd.objects.size(); i++)
    {
      if (const Sphere * s = dynamic_cast<const Sphere *>(sd.objects[i].get()))
      {
        const float c[3] = { s->center.x, s->center.y, s->center.z }, col[3] = { s->material.color.r, s->material.color.g, s->material.color.b };
        rfx_add_sphere(ctx, c, s->radius, s->material.type == Material::mtDielectric, col, s->material.reflectivity, s->material.transparency);
      }
      else if (const Triangle * t = dynamic_cast<const Triangle *>(sd.objects[i].get()))
      {
        const float v[9] = { t->v[0].x, t->v[0].y, t->v[0].z, t->v[1].x, t->v[1].y, t->v[1].z, t->v[2].x, t->v[2].y, t->v[2].z };
        const float col[3] = { t->material.color.r, t->material.color.g, t->material.color.b };
        const int idx = rfx_add_triangle(ctx, v, t->material.type == Material::mtDielectric, col, t->material.reflectivity, t->material.transparency);
        if (t->texture && idx >= 0)
          for (size_t k = 0; k < texPtr.size(); k++)
            if (texPtr[k] == t->texture) { rfx_set_triangle_texture(ctx, idx, int(k), t->uv); break; }
      }
    }
    uploadedScene = &sd;
    uploadedRevision = sd.revision;
    return true;
  }

public:
  Camera camera;
  Scene scene;
  unsigned int imageWidth, imageHeight;
  int additiveCounter;
  bool inProgress;

  Render(const char * exePath) : ctx(NULL), uploadedRevision(0), uploadedScene(NULL), cacheValid(false),
                                 imageWidth(0), imageHeight(0), additiveCounter(0), inProgress(false)
  {
    if (rfx_create(&ctx, deviceFromEnv()) != RFX_OK)
    {
      // no CPU fallback exists: a viewer without a B200 cannot render.  Say why, loudly, and leave a dead Render behind
      // (every call then returns false / black, the reference's own "bad state" convention).
      fprintf(stderr, "reflaxman_b200: %s\n", rfx_last_error(NULL));
      ctx = NULL;
    }
    else
    {
      // the reference seeds one LCG per translation unit from rand() during static initialisation (trace_math.h:34)
      const uint32_t s1 = uint32_t(rand()), s2 = uint32_t(rand());
      rfx_set_seeds(ctx, s1, s2);
    }
    scene.host = this;
    loadScene(exePath);
  }
  ~Render() { if (ctx) rfx_destroy(ctx); }
  Render(const Render &) = delete;
  Render & operator=(const Render &) = delete;

  void loadScene(const char * exePath)   // the reference's built-in demo scene, reference Render.cpp:25-55
  {
    const std::string dir(exePath);
    camera = Camera(Vector3(7.427f, 3.494f, -3.773f), Vector3(6.5981f, 3.127f, -3.352f), 1.05f);
    scene = Scene(Color(0.95f, 0.95f, 1.0f), 0.15f);
    scene.setSkyboxTexture((dir + "./textures/skybox.tga").c_str());
    scene.addLight(Vector3(11.8e9f, 4.26e9f, 3.08e9f), 3.48e8f, Color(1.0f, 1.0f, 0.95f), 0.85f);
    struct S { float x, y, z, r; Material::Type t; float cr, cg, cb, refl; };
    static const S spheres[] = {
      { -1.25f, 1.5f, -0.25f, 1.5f, Material::mtMetal, 1.0f, 1.0f, 1.0f, 1.0f },
      { 0.15f, 1.0f, 1.75f, 1.0f, Material::mtMetal, 1.0f, 1.0f, 1.0f, 0.95f },
      { -3.0f, 0.6f, -3.0f, 0.6f, Material::mtDielectric, 1.0f, 1.0f, 1.0f, 0.0f },
      { -0.5f, 0.5f, -2.5f, 0.5f, Material::mtDielectric, 0.5f, 1.0f, 0.15f, 0.75f },
      { 1.0f, 0.4f, -1.5f, 0.4f, Material::mtDielectric, 0.0f, 0.5f, 1.0f, 1.0f },
      { 1.8f, 0.4f, 0.1f, 0.4f, Material::mtMetal, 1.0f, 0.65f, 0.45f, 1.0f },
      { 1.7f, 0.5f, 1.9f, 0.5f, Material::mtMetal, 1.0f, 0.90f, 0.60f, 0.75f },
      { 0.6f, 0.6f, 4.2f, 0.6f, Material::mtMetal, 0.9f, 0.9f, 0.9f, 0.0f },
    };
    for (size_t i = 0; i < sizeof(spheres) / sizeof(spheres[0]); i++)
    {
      const S & s = spheres[i];
      scene.addSphere(Vector3(s.x, s.y, s.z), s.r, Material(s.t, Color(s.cr, s.cg, s.cb), s.refl, 0.0f));
    }
    Texture * floorTex = scene.addTexture((dir + "./textures/himiya.tga").c_str());
    const Material floorMat(Material::mtDielectric, Color(1.0f, 1.0f, 1.0f), 0.95f, 0.0f);
    const Vector3 a(-14.0f, 0.0f, -10.0f), b(-14.0f, 0.0f, 10.0f), c(14.0f, 0.0f, -10.0f), d(14.0f, 0.0f, 10.0f);
    scene.addTriangle(a, b, c, floorMat)->setTexture(floorTex, 0.0f, 0.0f, 0.0f, 1.0f, 1.0f, 0.0f);
    scene.addTriangle(b, d, c, floorMat)->setTexture(floorTex, 0.0f, 1.0f, 1.0f, 1.0f, 1.0f, 0.0f);
  }

  void setImageSize(unsigned int width, unsigned int height)   // reference Render.cpp:57-80
  {
    if (!ctx || !width || !height) return;
    if (rfx_set_image_size(ctx, width, height) != RFX_OK) return;
    imageWidth = width; imageHeight = height;
    additiveCounter = 0; inProgress = false; cacheValid = false;
  }

  void renderBegin(int reflectNum, int sampleNum, bool additive)   // reference Render.cpp:116-134
  {
    if (!ctx || !syncScene()) return;
    const float eye[3] = { camera.eye.x, camera.eye.y, camera.eye.z };
    rfx_set_camera(ctx, eye, &camera.view._11, camera.fov);
    if (rfx_render_begin(ctx, reflectNum, sampleNum, additive ? 1 : 0) != RFX_OK) return;
    additiveCounter = rfx_additive_counter(ctx);
    inProgress = true;
  }

  bool renderNext(unsigned int pixels)   // reference Render.cpp:136-215: true while the frame is incomplete
  {
    if (!ctx || !pixels || !inProgress) return false;
    const int rc = rfx_render_next(ctx, pixels);
    cacheValid = false;
    inProgress = rc == 1;
    return inProgress;
  }

  void renderAll(int reflectNum, int sampleNum, bool additive)   // reference Render.cpp:217-221 passes imageHeight as the pixel count
  {
    renderBegin(reflectNum, sampleNum, additive);
    renderNext(imageHeight);
  }

  float getRenderProgress() const { return ctx ? rfx_progress(ctx) : 0.0f; }

  Color imagePixel(int x, int y) const   // reference Render.cpp:103-114
  {
    if (!ctx || x < 0 || y < 0 || (unsigned)x >= imageWidth || (unsigned)y >= imageHeight) return Color(0, 0, 0);
    if (!cacheValid)
    {
      cache.resize(size_t(imageWidth) * imageHeight * 3);
      if (rfx_read_rgbf(ctx, cache.data()) != RFX_OK) return Color(0, 0, 0);
      cacheValid = true;
    }
    const float * p = &cache[(size_t(y) * imageWidth + x) * 3];
    return Color(p[0], p[1], p[2]);
  }

  void copyImage(Texture & texture) const   // reference Render.cpp:82-101: the RAW image (no additive divide) packed to ARGB
  {
    if (ctx && imageWidth == texture.getWidth() && imageHeight == texture.getHeight() && imageWidth && imageHeight)
      rfx_read_image(ctx, NULL, texture.getColorBuffer(), 0);
    else
      texture.clear(0);
  }

  rfx_ctx * context() { return ctx; }   // extension: the headless bench driver reaches the batch API through this

  // Scene::trace forwards here: n rays, one randDir each, in call order (reference Scene.cpp:73-75)
  bool traceRays(int n, const float * origins, const float * rays, int reflNumber, float * rgb)
  {
    return ctx && syncScene() && rfx_trace_rays(ctx, n, origins, rays, reflNumber, rgb) == RFX_OK;
  }
};

inline Color Scene::trace(Vector3 origin, Vector3 ray, int reflNumber) const
{
  float rgb[3] = { 0, 0, 0 };
  const float o[3] = { origin.x, origin.y, origin.z }, r[3] = { ray.x, ray.y, ray.z };
  if (host) host->traceRays(1, o, r, reflNumber, rgb);
  return Color(rgb[0], rgb[1], rgb[2]);
}
