/* TEST INFRASTRUCTURE ONLY -- oracle driver, not part of the product path.
 * Calls the UNMODIFIED algorithm templates of the reference (cholesky::cholinv::factor,
 * cholinv.hpp:6-28) on the reference's own generator (structure.hpp:69-103) and validator
 * (test/cholesky/validate.hpp:7-49); protocol mirrors bench/cholesky/cholinv.cpp:38-60
 * (warm-up factor, barrier, MPI_Wtime around factor) with the residual block re-enabled.
 *
 * usage: ref_cholinv n complete_inv split bc_mult policy iters [dump_prefix]
 *   policy: 0 ReplicateCommComp (default for P>1), 1 ReplicateComp, 2 NoReplication (benchmarked, wrong at P>1)
 *   ranks: env MINIMPI_NP (1 or a cube: 8, 27); c = cbrt(P) as in the bench (:34-35)
 * prints one JSON line on rank 0; with dump_prefix every rank writes A / R / Rinv local blocks.
 */
#include "src/alg/cholesky/cholinv/cholinv.h"
#include "test/cholesky/validate.h"
#include <string>

using T = double; using U = int64_t; using MatrixType = matrix<T,U,rect>;

static void dump(const std::string& path, const T* p, U n){ FILE* f = fopen(path.c_str(),"wb"); fwrite(p,sizeof(T),(size_t)n,f); fclose(f); }

template<class CholType>
static int run(int argc, char** argv, int rank, int size){
  U n = atol(argv[1]); bool complete_inv = atoi(argv[2]); U split = atoi(argv[3]); U bcm = atoi(argv[4]);
  int iters = atoi(argv[6]); std::string prefix = argc > 7 ? argv[7] : "";
  size_t c = std::nearbyint(std::ceil(pow(size,1./3.)));
  auto topo = topo::square(MPI_COMM_WORLD,c,0,0);
  MatrixType A(n,n,topo.d,topo.d);
  A.distribute_symmetric(topo.x,topo.y,topo.d,topo.d,rank/topo.c,true);
  typename CholType::template info<T,U> pack(complete_inv,split,bcm,'U');
  CholType::factor(A,pack,topo);
  double best = 1e30, sum = 0;
  for (int i=0;i<iters;i++){
    MPI_Barrier(MPI_COMM_WORLD);
    double t0 = MPI_Wtime();
    CholType::factor(A,pack,topo);
    double t = MPI_Wtime()-t0;
    MPI_Allreduce(MPI_IN_PLACE,&t,1,MPI_DOUBLE,MPI_MAX,MPI_COMM_WORLD);
    best = std::min(best,t); sum += t;
  }
  if (!prefix.empty()){
    std::string r = std::to_string(rank);
    dump(prefix+".A."+r+".bin",A.data(),A.num_elems());
    dump(prefix+".R."+r+".bin",pack.R.data(),pack.R.num_elems());
    dump(prefix+".Rinv."+r+".bin",pack.Rinv.data(),pack.Rinv.num_elems());
  }
  double res_local = cholesky::validate<CholType>::residual(A,pack,topo), res = 0;
  MPI_Reduce(&res_local,&res,1,MPI_DOUBLE,MPI_MAX,0,MPI_COMM_WORLD);
  if (rank==0){
    printf("{\"alg\":\"cholinv\",\"n\":%ld,\"P\":%d,\"c\":%zu,\"d\":%zu,\"complete_inv\":%d,\"split\":%ld,\"bc_mult_dim\":%ld,\"bc_dim\":%ld,"
           "\"iters\":%d,\"time_best_s\":%.6f,\"time_mean_s\":%.6f,\"residual\":%.6e}\n",
           (long)n,size,(size_t)topo.c,(size_t)topo.d,(int)complete_inv,(long)split,(long)bcm,(long)pack.bcDimension,iters,best,iters?sum/iters:0.,res);
    fflush(stdout);
  }
  return 0;
}

int main(int argc, char** argv){
  if (argc < 7){ fprintf(stderr,"usage: %s n complete_inv split bc_mult policy iters [dump_prefix]\n",argv[0]); return 2; }
  int rank,size,provided; MPI_Init_thread(&argc,&argv,MPI_THREAD_SINGLE,&provided);
  MPI_Comm_rank(MPI_COMM_WORLD,&rank); MPI_Comm_size(MPI_COMM_WORLD,&size);
  using namespace cholesky; namespace pc = cholesky::policy::cholinv;
  int policy = atoi(argv[5]); int rc;
  if (policy==0)      rc = run<cholinv<pc::Serialize,pc::SaveIntermediates,pc::ReplicateCommComp>>(argc,argv,rank,size);
  else if (policy==1) rc = run<cholinv<pc::Serialize,pc::SaveIntermediates,pc::ReplicateComp>>(argc,argv,rank,size);
  else                rc = run<cholinv<pc::Serialize,pc::SaveIntermediates,pc::NoReplication>>(argc,argv,rank,size);
  MPI_Finalize();
  return rc;
}
