/* TEST INFRASTRUCTURE ONLY -- oracle driver, not part of the product path.
 * Calls the reference's qr::cacqr::factor (cacqr.hpp:217-248) on its own tall-skinny generator
 * (structure.hpp:106-129, key = rank/c as in bench/qr/cacqr.cpp:34) and validators
 * (test/qr/validate.hpp:7-52); timing protocol of bench/qr/cacqr.cpp:47-53.
 *
 * usage: ref_cacqr variant m n c complete_inv split bc_mult iters [dump_prefix]
 *   ranks: env MINIMPI_NP = c*c*d ; c=1 -> 1D (d = P), c=d -> 3D. cholinv policy = ReplicateCommComp.
 */
#include "src/alg/qr/cacqr/cacqr.h"
#include "test/qr/validate.h"
#include <string>

using T = double; using U = int64_t; using MatrixType = matrix<T,U,rect>;
static void dump(const std::string& path, const T* p, U n){ FILE* f = fopen(path.c_str(),"wb"); fwrite(p,sizeof(T),(size_t)n,f); fclose(f); }

int main(int argc, char** argv){
  if (argc < 9){ fprintf(stderr,"usage: %s variant m n c complete_inv split bc_mult iters [dump_prefix]\n",argv[0]); return 2; }
  int rank,size,provided; MPI_Init_thread(&argc,&argv,MPI_THREAD_SINGLE,&provided);
  MPI_Comm_rank(MPI_COMM_WORLD,&rank); MPI_Comm_size(MPI_COMM_WORLD,&size);
  size_t variant = atoi(argv[1]); U m = atol(argv[2]); U n = atol(argv[3]); U c = atoi(argv[4]);
  bool complete_inv = atoi(argv[5]); U split = atoi(argv[6]); U bcm = atoi(argv[7]); int iters = atoi(argv[8]);
  std::string prefix = argc > 9 ? argv[9] : "";
  namespace pc = cholesky::policy::cholinv;
  using ci_type = cholesky::cholinv<pc::Serialize,pc::SaveIntermediates,pc::ReplicateCommComp>;
  using qr_type = qr::cacqr<qr::policy::cacqr::Serialize,qr::policy::cacqr::SaveIntermediates>;
  {
    auto RectTopo = topo::rect(MPI_COMM_WORLD,c,0,0);
    MatrixType A(n,m,RectTopo.c,RectTopo.d);
    A.distribute_random(RectTopo.x,RectTopo.y,RectTopo.c,RectTopo.d,rank/RectTopo.c);
    ci_type::info<T,U> ci_pack(complete_inv,split,bcm,'U');
    qr_type::info<T,U,ci_type> pack(variant,ci_pack);
    qr_type::factor(A,pack,RectTopo);
    double best = 1e30, sum = 0;
    for (int i=0;i<iters;i++){
      MPI_Barrier(MPI_COMM_WORLD);
      double t0 = MPI_Wtime();
      qr_type::factor(A,pack,RectTopo);
      double t = MPI_Wtime()-t0;
      MPI_Allreduce(MPI_IN_PLACE,&t,1,MPI_DOUBLE,MPI_MAX,MPI_COMM_WORLD);
      best = std::min(best,t); sum += t;
    }
    if (!prefix.empty()){
      std::string r = std::to_string(rank);
      dump(prefix+".A."+r+".bin",A.data(),A.num_elems());
      dump(prefix+".Q."+r+".bin",pack.Q.data(),pack.Q.num_elems());
      dump(prefix+".R."+r+".bin",pack.R.data(),pack.R.num_elems());
    }
    double res_l = qr::validate<qr_type>::residual(A,pack,RectTopo), res = 0;
    double orth_l = qr::validate<qr_type>::orthogonality(A,pack,RectTopo), orth = 0;
    MPI_Reduce(&res_l,&res,1,MPI_DOUBLE,MPI_MAX,0,MPI_COMM_WORLD);
    MPI_Reduce(&orth_l,&orth,1,MPI_DOUBLE,MPI_MAX,0,MPI_COMM_WORLD);
    if (rank==0){
      printf("{\"alg\":\"cacqr\",\"variant\":%zu,\"m\":%ld,\"n\":%ld,\"P\":%d,\"c\":%ld,\"d\":%zu,\"iters\":%d,"
             "\"time_best_s\":%.6f,\"time_mean_s\":%.6f,\"residual\":%.6e,\"orthogonality\":%.6e}\n",
             variant,(long)m,(long)n,size,(long)c,(size_t)RectTopo.d,iters,best,iters?sum/iters:0.,res,orth);
      fflush(stdout);
    }
  }
  MPI_Finalize();
  return 0;
}
