// kmeans driver -- CLI of the reference's drivers/kmeans.cpp:192-197 / drivers/in_mem_kmeans.cpp:155-158:
//   <points> <centers> <npoints> <ndims> <ncenters> [n_iters = 1]
// The reference runs one Lloyd iteration (drivers/kmeans.cpp:219) and leaves the new centers in the
// centers file; the optional sixth argument runs more (BASELINE.json cfg-5 uses 20).
#include "driver_common.h"

int main(int argc, char** argv) {
  if (argc != 6 && argc != 7) drv::usage_exit("kmeans <points> <centers> <npoints> <ndims> <ncenters> [n_iters]");
  flash::flash_setup("/tmp/");
  const FBLAS_UINT npoints = drv::to_u(argv[3]), ndims = drv::to_u(argv[4]), ncenters = drv::to_u(argv[5]);
  const FBLAS_UINT iters = argc == 7 ? drv::to_u(argv[6]) : 1;
  auto points = flash::map_file<FPTYPE>(argv[1], flash::Mode::READ);
  auto centers = flash::map_file<FPTYPE>(argv[2], flash::Mode::READWRITE);
  drv::StopWatch sw;
  const FBLAS_INT rc = flash::kmeans_lloyd(points, centers, npoints, ndims, ncenters, iters);
  drv::report("kmeans_lloyd()", sw.seconds(), rc);
  flash::unmap_file(points);
  flash::unmap_file(centers);
  flash::flash_destroy();
  return rc == 0 ? 0 : 1;
}
