// CPU-only check of the shim's on-disk formats (SURVEY f-4): Texture save (32-bpp BMP / TGA) and TGA load.
#include "Texture.h"
#include <cstdio>
int main(int argc, char ** argv)
{
  if (argc < 2) return 2;
  const std::string dir = argv[1];
  Texture t(5, 3);
  for (unsigned i = 0; i < 15; i++) t.getColorBuffer()[i] = 0x01020300u * (i + 1) + i;
  if (!t.saveToFile((dir + "/a.bmp").c_str())) return 3;
  if (!t.saveToFile((dir + "/a.tga").c_str())) return 4;
  if (t.saveToFile((dir + "/a.png").c_str())) return 5;        // unknown extension -> false (reference Texture.cpp:190-205)
  Texture u((dir + "/a.tga").c_str());
  if (u.getWidth() != 5 || u.getHeight() != 3) return 6;
  for (unsigned i = 0; i < 15; i++) if (u.getColorBuffer()[i] != t.getColorBuffer()[i]) return 7;
  Texture v;
  if (v.loadFromFile((dir + "/missing.tga").c_str()) || v.getWidth() != 0 || !v.empty()) return 8;   // failed load -> empty texture
  if (v.loadFromFile((dir + "/a.bmp").c_str())) return 9;      // only .tga loads (reference Texture.cpp:175-188)
  printf("ok\n");
  return 0;
}
