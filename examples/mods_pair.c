/* mods_pair.c -- the smallest host program over the C ABI (include/modsgpu.h): one image pair through the deep
 * configuration (Hessian -> AffNet -> OriNet -> HardNet++ -> FGINN -> duplicate filter -> LO-RANSAC(H)), the work one
 * iteration of the reference's mods.cpp:202-356 does for iters_HessianZMQ.ini.  Plain C on purpose: nothing of the ABI
 * needs C++.  Images are binary PPM (P6) or PGM (P5); PNG/JPEG decoding is the reference's own business (cv::imread).
 *
 *   make example
 *   examples/mods_pair img1.ppm img2.ppm weights/ [overlap] > correspondences.txt
 *
 * Output: `x1 y1 x2 y2` per verified correspondence (reproj_kp of both regions), the counts and H on stderr. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "modsgpu.h"

static int skip_ws_and_comments(FILE* f) {
  int c;
  while ((c = fgetc(f)) != EOF) {
    if (c == '#') { while ((c = fgetc(f)) != EOF && c != '\n') {} continue; }
    if (c != ' ' && c != '\t' && c != '\n' && c != '\r') { ungetc(c, f); return 0; }
  }
  return -1;
}

/* P5 / P6 with maxval 255 -> interleaved BGR (cv::imread layout); returns NULL on error */
static uint8_t* read_pnm_bgr(const char* path, int* w, int* h) {
  FILE* f = fopen(path, "rb");
  if (!f) { perror(path); return NULL; }
  char magic[3] = {0, 0, 0};
  int maxval = 0;
  uint8_t* bgr = NULL;
  if (fread(magic, 1, 2, f) == 2 && magic[0] == 'P' && (magic[1] == '5' || magic[1] == '6') && !skip_ws_and_comments(f) &&
      fscanf(f, "%d", w) == 1 && !skip_ws_and_comments(f) && fscanf(f, "%d", h) == 1 && !skip_ws_and_comments(f) &&
      fscanf(f, "%d", &maxval) == 1 && maxval == 255 && *w > 0 && *h > 0 && fgetc(f) != EOF) {
    const size_t n = (size_t)*w * *h, ch = magic[1] == '6' ? 3 : 1;
    uint8_t* raw = (uint8_t*)malloc(n * ch);
    bgr = (uint8_t*)malloc(n * 3);
    if (raw && bgr && fread(raw, ch, n, f) == n) {
      for (size_t i = 0; i < n; i++) {
        if (ch == 3) { bgr[3 * i] = raw[3 * i + 2]; bgr[3 * i + 1] = raw[3 * i + 1]; bgr[3 * i + 2] = raw[3 * i]; }
        else bgr[3 * i] = bgr[3 * i + 1] = bgr[3 * i + 2] = raw[i];
      }
    } else { free(bgr); bgr = NULL; }
    free(raw);
  }
  if (!bgr) fprintf(stderr, "%s: not a binary PGM / PPM with maxval 255\n", path);
  fclose(f);
  return bgr;
}

int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage: %s img1.ppm img2.ppm weights_dir [overlap]\n", argv[0]); return 2; }
  int w1, h1, w2, h2;
  uint8_t* a = read_pnm_bgr(argv[1], &w1, &h1);
  uint8_t* b = read_pnm_bgr(argv[2], &w2, &h2);
  if (!a || !b) return 1;
  if (w1 != w2 || h1 != h2) { fprintf(stderr, "modsgpu_pair_pipeline takes two images of one size; use modsgpu_image_from_bgr8 + modsgpu_pair_pipeline_images otherwise\n"); return 1; }
  modsgpu_ctx* ctx = NULL;
  int rc = modsgpu_create(0, &ctx);
  if (rc) { fprintf(stderr, "modsgpu_create: error %d (an sm_100 device is required; there is no CPU path)\n", rc); return 1; }
  const char* names[3] = {"affnet.npz", "orinet.npz", "hardnet.npz"};
  const modsgpu_net nets[3] = {MODSGPU_AFFNET, MODSGPU_ORINET, MODSGPU_HARDNET};
  for (int i = 0; i < 3 && !rc; i++) {
    char path[4096];
    snprintf(path, sizeof(path), "%s/%s", argv[3], names[i]);
    rc = modsgpu_load_weights(ctx, nets[i], path);
  }
  if (!rc && argc > 4 && atoi(argv[4]) != 0) rc = modsgpu_set_pair_overlap(ctx, 1);   /* the two images side by side */
  enum { CAP = 1 << 16 };
  double* xy = (double*)malloc(sizeof(double) * 4 * CAP);
  modsgpu_pair_result res;
  memset(&res, 0, sizeof(res));
  if (!rc) rc = modsgpu_pair_pipeline(ctx, a, b, w1, h1, 12345ull /* the reference seeds with time(NULL) */, &res, xy, CAP);
  if (rc) {
    fprintf(stderr, "error %d: %s\n", rc, modsgpu_last_error(ctx));
  } else {
    fprintf(stderr, "keypoints %d / %d, regions %d / %d, descriptors %d / %d, tentatives %d, unique %d, verified %d\n", res.keypoints[0],
            res.keypoints[1], res.regions[0], res.regions[1], res.descriptors[0], res.descriptors[1], res.tentatives,
            res.unique_tentatives, res.inliers);
    fprintf(stderr, "H = [%.9g %.9g %.9g; %.9g %.9g %.9g; %.9g %.9g %.9g]\n", res.H[0], res.H[1], res.H[2], res.H[3], res.H[4],
            res.H[5], res.H[6], res.H[7], res.H[8]);
    for (int i = 0; i < res.inliers && i < CAP; i++) printf("%.6f %.6f %.6f %.6f\n", xy[4 * i], xy[4 * i + 1], xy[4 * i + 2], xy[4 * i + 3]);
  }
  free(xy); free(a); free(b);
  modsgpu_destroy(ctx);
  return rc ? 1 : 0;
}
