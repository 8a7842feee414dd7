// Mutation fuzzer for the asset readers (JPEG, PNG, TGA, OBJ), meant to be built with the sanitizers:
//   g++ -std=c++17 -O1 -g -fsanitize=address,undefined -Ics397raytracingsp22_b200/csrc -I/usr/local/cuda/include \
//       tools/fuzz_decoders.cpp cs397raytracingsp22_b200/csrc/rt_{jpeg,png,lower}.cpp -o build/fuzz_decoders
//   build/fuzz_decoders 8000 tests/golden/jpeg/*.jpg assets/texture/*.png some.tga some.obj
// Every mutated input must either decode or be refused; any sanitizer report is a bug.  (Round 1: 56 k JPEG and 40 k
// PNG/TGA/OBJ inputs clean after one fix - Huffman tables that over-subscribe the code space are now refused.)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>
#include "rt_lower.h"
static std::vector<uint8_t> slurp(const char* p){ FILE* f=fopen(p,"rb"); std::vector<uint8_t> v; if(!f) return v; fseek(f,0,SEEK_END); long n=ftell(f); fseek(f,0,SEEK_SET); v.resize(n); if(fread(v.data(),1,n,f)!=(size_t)n) v.clear(); fclose(f); return v; }
int main(int argc,char**argv){
  int iters = atoi(argv[1]); std::mt19937 rng(12345); long ok=0, bad=0;
  for(int a=2;a<argc;++a){
    std::vector<uint8_t> base=slurp(argv[a]); std::string name=argv[a];
    int kind = name.size()>4 && name.substr(name.size()-4)==".png" ? 1 : (name.substr(name.size()-4)==".tga" ? 2 : (name.substr(name.size()-4)==".obj" ? 3 : 0));
    for(int it=0; it<iters; ++it){
      std::vector<uint8_t> d=base;
      int nm = 1 + rng()%8;
      for(int m=0;m<nm;++m){
        int op=rng()%4; size_t pos=rng()%d.size();
        if(op==0) d[pos]=(uint8_t)rng();
        else if(op==1) d[pos]^= (uint8_t)(1u<<(rng()%8));
        else if(op==2 && d.size()>16) d.resize(16 + rng()%(d.size()-16));
        else if(op==3) { size_t n = 1 + rng()%16; for(size_t k=0;k<n && pos+k<d.size();++k) d[pos+k]=0xFF; }
        if (d.empty()) d.push_back(0);
      }
      uint8_t* rgb=nullptr; uint32_t w=0,h=0; std::string err; int rc;
      if(kind==0) rc=rt::jpeg_decode(d.data(), d.size(), &rgb,&w,&h,err);
      else if(kind==1) rc=rt::png_decode(d.data(), d.size(), &rgb,&w,&h,err);
      else if(kind==2) rc=rt::tga_decode(d.data(), d.size(), &rgb,&w,&h,err);
      else { rt_obj_mesh m; memset(&m,0,sizeof m); rc=rt::obj_parse((const char*)d.data(), d.size(), &m, err); if(rc==0){ free(m.pos); free(m.nrm); free(m.uv); free(m.idx);} }
      if(rc==0){ ++ok; if(rgb){ volatile uint8_t s=0; for(size_t i=0;i<(size_t)w*h*3;i+=97) s+=rgb[i]; free(rgb);} } else ++bad;
    }
    printf("%s: ok %ld bad %ld\n", argv[a], ok, bad); fflush(stdout);
  }
  return 0;
}
