// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// Headless driver around the UNMODIFIED ReflaxMan sources.  It is compiled by oracle/Makefile
// together with /root/reference/src/common/*.cpp (sources stay where they lie; nothing is copied
// into this repository) and the outputs land in oracle/_ref/ (git-ignored).  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may execute it.
//
// It uses nothing but the reference's public API, exactly as Pulse's screenshot flow does
// (reference Pulse.cpp:174-176 setImageSize+renderBegin, :186 renderNext, :206-208 copyImage):
//   Render r(texDir); r.setImageSize(W,H); r.renderBegin(refl,samples,additive);
//   while (r.renderNext(chunk)) {}   r.imagePixel(x,y) / r.copyImage(tex)
// Render::camera and Render::scene are public members (reference Render.h:22-23), so custom
// cameras and scenes need no source change.
//
// Seed pinning: every reference TU that includes trace_math.h owns `static int g_seed = rand();`
// (reference trace_math.h:34).  We interpose rand() so both live streams (Vector3.cpp: randDir,
// Render.cpp: additive jitter) start from $RFX_SEED (default 12345) regardless of link order.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <chrono>
#include <math.h>

#include "Render.h"

extern "C" int rand(void) noexcept
{
  const char * s = getenv("RFX_SEED");
  return s ? (int)strtol(s, NULL, 0) : 12345;
}

namespace
{
  struct Cam { float eye[3]; float view[9]; float fov; bool lookat; float at[3]; };

  double now()
  {
    using namespace std::chrono;
    return duration<double>(steady_clock::now().time_since_epoch()).count();
  }

  void die(const char * msg)
  {
    fprintf(stderr, "ref_harness: %s\n", msg);
    exit(2);
  }

  Material::Type mtype(int t) { return t ? Material::mtDielectric : Material::mtMetal; }

  // scene text format (ours, see reflaxman_b200/scenes.py): one record per line
  void loadSceneFile(Render & r, const char * path)
  {
    FILE * f = fopen(path, "r");
    if (!f) die("cannot open scene file");
    char tag[64];
    std::vector<Texture*> textures;
    bool haveAmbient = false;

    while (fscanf(f, "%63s", tag) == 1)
    {
      if (!strcmp(tag, "ambient"))
      {
        float c[3], p;
        if (fscanf(f, "%f %f %f %f", &c[0], &c[1], &c[2], &p) != 4) die("bad ambient");
        r.scene = Scene(Color(c[0], c[1], c[2]), p);  // drops the default objects (leaks them; harmless)
        haveAmbient = true;
      }
      else if (!haveAmbient) die("scene file must start with 'ambient'");
      else if (!strcmp(tag, "skybox"))
      {
        char p[1024];
        if (fscanf(f, "%1023s", p) != 1) die("bad skybox");
        r.scene.setSkyboxTexture(p);  // "-" fails to load -> checker fallback
      }
      else if (!strcmp(tag, "light"))
      {
        float o[3], rad, c[3], p;
        if (fscanf(f, "%f %f %f %f %f %f %f %f", &o[0], &o[1], &o[2], &rad, &c[0], &c[1], &c[2], &p) != 8) die("bad light");
        r.scene.addLight(Vector3(o[0], o[1], o[2]), rad, Color(c[0], c[1], c[2]), p);
      }
      else if (!strcmp(tag, "texture"))
      {
        char p[1024];
        if (fscanf(f, "%1023s", p) != 1) die("bad texture");
        textures.push_back(r.scene.addTexture(p));
      }
      else if (!strcmp(tag, "sphere"))
      {
        float c[3], rad, col[3], refl, transp; int t;
        if (fscanf(f, "%f %f %f %f %d %f %f %f %f %f", &c[0], &c[1], &c[2], &rad, &t, &col[0], &col[1], &col[2], &refl, &transp) != 10) die("bad sphere");
        r.scene.addSphere(Vector3(c[0], c[1], c[2]), rad, Material(mtype(t), Color(col[0], col[1], col[2]), refl, transp));
      }
      else if (!strcmp(tag, "tri"))
      {
        float v[9], col[3], refl, transp, uv[6]; int t, tex;
        for (int i = 0; i < 9; i++) if (fscanf(f, "%f", &v[i]) != 1) die("bad tri");
        if (fscanf(f, "%d %f %f %f %f %f %d", &t, &col[0], &col[1], &col[2], &refl, &transp, &tex) != 7) die("bad tri");
        for (int i = 0; i < 6; i++) if (fscanf(f, "%f", &uv[i]) != 1) die("bad tri uv");
        Triangle * tr = r.scene.addTriangle(Vector3(v[0], v[1], v[2]), Vector3(v[3], v[4], v[5]), Vector3(v[6], v[7], v[8]),
                                            Material(mtype(t), Color(col[0], col[1], col[2]), refl, transp));
        if (tex >= 0)
        {
          if (tex >= (int)textures.size()) die("tri references unknown texture");
          tr->setTexture(textures[tex], uv[0], uv[1], uv[2], uv[3], uv[4], uv[5]);
        }
      }
      else die("unknown scene record");
    }
    fclose(f);
  }

  std::vector<Cam> loadCamFile(const char * path)
  {
    std::vector<Cam> cams;
    FILE * f = fopen(path, "r");
    if (!f) die("cannot open camera file");
    char tag[32];
    while (fscanf(f, "%31s", tag) == 1)
    {
      Cam c; memset(&c, 0, sizeof(c));
      if (!strcmp(tag, "lookat"))
      {
        c.lookat = true;
        if (fscanf(f, "%f %f %f %f %f %f %f", &c.eye[0], &c.eye[1], &c.eye[2], &c.at[0], &c.at[1], &c.at[2], &c.fov) != 7) die("bad lookat");
      }
      else if (!strcmp(tag, "view"))
      {
        c.lookat = false;
        for (int i = 0; i < 3; i++) if (fscanf(f, "%f", &c.eye[i]) != 1) die("bad view");
        for (int i = 0; i < 9; i++) if (fscanf(f, "%f", &c.view[i]) != 1) die("bad view");
        if (fscanf(f, "%f", &c.fov) != 1) die("bad view");
      }
      else die("unknown camera record");
      cams.push_back(c);
    }
    fclose(f);
    return cams;
  }

  void applyCam(Render & r, const Cam & c)
  {
    if (c.lookat)
      r.camera = Camera(Vector3(c.eye[0], c.eye[1], c.eye[2]), Vector3(c.at[0], c.at[1], c.at[2]), c.fov);
    else
    {
      r.camera.eye = Vector3(c.eye[0], c.eye[1], c.eye[2]);
      r.camera.view = Matrix33(c.view[0], c.view[1], c.view[2], c.view[3], c.view[4], c.view[5], c.view[6], c.view[7], c.view[8]);
      r.camera.fov = c.fov;
    }
  }

  void printCam(const Render & r)
  {
    const Matrix33 & m = r.camera.view;
    printf("\"eye\": [\"%a\", \"%a\", \"%a\"], \"view\": [\"%a\", \"%a\", \"%a\", \"%a\", \"%a\", \"%a\", \"%a\", \"%a\", \"%a\"], \"fov\": \"%a\"",
      r.camera.eye.x, r.camera.eye.y, r.camera.eye.z, m._11, m._12, m._13, m._21, m._22, m._23, m._31, m._32, m._33, r.camera.fov);
  }
}

int main(int argc, char ** argv)
{
  unsigned W = 1024, H = 768;
  int refl = 20, samples = 1, passes = 1, frames = 1;
  bool additive = false;
  unsigned chunk = 0;
  int rowsMod = 0, rowsRem = 0;
  const char * sceneFile = NULL, * camFile = NULL, * outPath = NULL;
  std::string texDir = "/nonexistent/";
  std::vector<int> dump;  // frame indices to dump (empty = all)

  for (int i = 1; i < argc; i++)
  {
    std::string a = argv[i];
    #define NEED(n) if (i + (n) >= argc) die("missing argument value")
    if (a == "--size") { NEED(2); W = atoi(argv[++i]); H = atoi(argv[++i]); }
    else if (a == "--refl") { NEED(1); refl = atoi(argv[++i]); }
    else if (a == "--samples") { NEED(1); samples = atoi(argv[++i]); }
    else if (a == "--additive") { NEED(1); additive = true; passes = atoi(argv[++i]); }
    else if (a == "--frames") { NEED(1); frames = atoi(argv[++i]); }
    else if (a == "--chunk") { NEED(1); chunk = (unsigned)strtoul(argv[++i], NULL, 0); }
    else if (a == "--rows") { NEED(2); rowsMod = atoi(argv[++i]); rowsRem = atoi(argv[++i]); }
    else if (a == "--scene") { NEED(1); sceneFile = argv[++i]; }
    else if (a == "--cams") { NEED(1); camFile = argv[++i]; }
    else if (a == "--texdir") { NEED(1); texDir = argv[++i]; }
    else if (a == "--out") { NEED(1); outPath = argv[++i]; }
    else if (a == "--dump") { NEED(1); dump.push_back(atoi(argv[++i])); }
    else die("unknown option");
    #undef NEED
  }
  if (!chunk) chunk = W * H;

  Render r(texDir.c_str());   // default scene (reference Render.cpp:25-55); textures resolve to texDir + "./textures/*.tga"
  if (sceneFile) loadSceneFile(r, sceneFile);
  std::vector<Cam> cams;
  if (camFile) cams = loadCamFile(camFile);

  r.setImageSize(W, H);
  FILE * out = outPath ? fopen(outPath, "wb") : NULL;
  if (outPath && !out) die("cannot open output");

  printf("{\"width\": %u, \"height\": %u, \"refl\": %d, \"samples\": %d, \"additive\": %s, \"passes\": %d, \"frames\": [\n",
         W, H, refl, samples, additive ? "true" : "false", passes);

  double total = 0;
  for (int f = 0; f < frames; f++)
  {
    if (!cams.empty()) applyCam(r, cams[f % cams.size()]);
    double t0 = now();

    if (rowsMod > 0)
    {
      // harness-parallel throughput mode: the reference has no threads, so each worker process traces an interleaved
      // set of rows through the reference's public Scene::trace, generating rays exactly as reference Render.cpp:146-156.
      const Vector3 origin = r.camera.eye;
      const float rz = float(W) / 2.0f / tanf(r.camera.fov / 2.0f);
      const float wh = W / 2.0f, hh = H / 2.0f;
      std::vector<Color> img(size_t(W) * H, Color(0, 0, 0));
      for (unsigned y = (unsigned)rowsRem; y < H; y += (unsigned)rowsMod)
        for (unsigned x = 0; x < W; x++)
        {
          Vector3 ray(float(x) - wh, float(y) - hh, rz);
          ray = r.camera.view * ray;
          img[x + size_t(y) * W] = r.scene.trace(origin, ray, refl);
        }
      double t1 = now();
      total += t1 - t0;
      printf("%s{\"frame\": %d, \"seconds\": %.6f, \"rows_mod\": %d, \"rows_rem\": %d, ", f ? ",\n" : "", f, t1 - t0, rowsMod, rowsRem);
      printCam(r);
      printf("}");
      if (out)
      {
        std::vector<float> rgb(size_t(W) * H * 3);
        for (size_t p = 0; p < size_t(W) * H; p++) { rgb[3 * p] = img[p].r; rgb[3 * p + 1] = img[p].g; rgb[3 * p + 2] = img[p].b; }
        fwrite(rgb.data(), sizeof(float), rgb.size(), out);
        std::vector<uint32_t> px(size_t(W) * H);
        for (size_t p = 0; p < size_t(W) * H; p++) px[p] = img[p].argb();
        fwrite(px.data(), 4, px.size(), out);
      }
      continue;
    }

    for (int p = 0; p < passes; p++)
    {
      r.renderBegin(refl, samples, additive);
      while (r.renderNext(chunk)) {}
    }
    double t1 = now();
    total += t1 - t0;
    printf("%s{\"frame\": %d, \"seconds\": %.6f, \"additiveCounter\": %d, ", f ? ",\n" : "", f, t1 - t0, r.additiveCounter);
    printCam(r);
    printf("}");

    bool want = dump.empty();
    for (size_t k = 0; k < dump.size(); k++) if (dump[k] == f) want = true;
    if (out && want)
    {
      std::vector<float> rgb(size_t(W) * H * 3);
      for (unsigned y = 0; y < H; y++)
        for (unsigned x = 0; x < W; x++)
        {
          const Color c = r.imagePixel(x, y);
          float * d = &rgb[(size_t(y) * W + x) * 3];
          d[0] = c.r; d[1] = c.g; d[2] = c.b;
        }
      fwrite(rgb.data(), sizeof(float), rgb.size(), out);
      // ARGB through the front ends' own path: imagePixel(x,y).argb() (reference Pulse.cpp:455-458)
      std::vector<uint32_t> px(size_t(W) * H);
      for (unsigned y = 0; y < H; y++)
        for (unsigned x = 0; x < W; x++)
          px[size_t(y) * W + x] = r.imagePixel(x, y).argb();
      fwrite(px.data(), 4, px.size(), out);
    }
    if (additive && f + 1 < frames) r.setImageSize(W, H);  // restart accumulation for the next frame
  }
  printf("\n], \"total_seconds\": %.6f}\n", total);
  if (out) fclose(out);
  return 0;
}
