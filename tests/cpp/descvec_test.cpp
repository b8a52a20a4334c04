// CPU check of the host mirror's descriptor container (csrc/host/mods_host.h: DescVec): the std::vector<float> surface the
// reference's call sites use (size / data / operator[] / assign), cheap copies that share one block, views into a
// describe call's read-back block, and region lists that survive the block's original owner.
#include "../../mods_light_zmq_b200/csrc/host/mods_host.h"
#include <cstdio>
#include <numeric>

using namespace modsb200;

#define CHECK(c) do { if (!(c)) { std::printf("FAILED %s:%d %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

int main() {
  AffineRegion r;
  CHECK(r.desc.empty() && r.desc.size() == 0 && r.desc.data() == nullptr);
  std::vector<float> v(128);
  std::iota(v.begin(), v.end(), 0.f);
  r.desc.assign(v.begin(), v.end());
  CHECK(r.desc.size() == 128 && r.desc[5] == 5.f && r.desc.data()[127] == 127.f);
  AffineRegion c = r;                                  // a copy shares the block
  CHECK(c.desc.data() == r.desc.data() && c.desc.block() == r.desc.block());
  r.desc.assign(v.begin(), v.begin() + 4);             // re-assigning one does not touch the other
  CHECK(r.desc.size() == 4 && c.desc.size() == 128 && c.desc[100] == 100.f);
  AffineRegionVector list(3);
  {
    auto blk = std::make_shared<const std::vector<float>>(3 * 128, 7.f);
    for (int i = 0; i < 3; i++) list[i].desc.view(blk, (size_t)i * 128, 128);
  }                                                    // the views keep the block alive
  CHECK(list[2].desc.size() == 128 && list[2].desc[127] == 7.f);
  CHECK(list[1].desc.data() == list[0].desc.data() + 128 && list[1].desc.offset() == 128);
  TentativeCorrespExt t;
  t.first = list[0]; t.second = list[2];
  std::vector<TentativeCorrespExt> tl(1000, t);        // what the tentative lists do: copies without allocations
  CHECK(tl[999].second.desc.data() == list[2].desc.data());
  list.clear();
  CHECK(tl[0].first.desc[0] == 7.f);
  std::printf("descvec ok\n");
  return 0;
}
