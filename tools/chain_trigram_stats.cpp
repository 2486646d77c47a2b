// One-off measurement behind DESIGN.md 7b: of the hash-chain entries a level-6 walk visits (<= 128 hops, distance < 32506), how
// many share the position's first two bytes (= its trigram, the hash being equal)?  g++ -O2 -o tri tools/chain_trigram_stats.cpp; ./tri buf.bin ...
// Bench workload (8 classes x 256 KiB): 95.0 % -- skipping hash colliders exactly would save nothing.
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cstring>
using namespace std;
int main(int argc,char**argv){
  long tot_h=0,tot_t=0,tot_first=0,npos=0,nfirst=0; long walk_t=0;
  for(int a=1;a<argc;a++){
    FILE*f=fopen(argv[a],"rb"); fseek(f,0,SEEK_END); long n=ftell(f); fseek(f,0,SEEK_SET); vector<uint8_t>d(n+8); fread(d.data(),1,n,f); fclose(f);
    vector<int> head(32768,-1), prev(n,-1);
    for(long p=0;p+2<n;p++){
      int h=((d[p]<<10)^(d[p+1]<<5)^d[p+2])&0x7FFF;
      prev[p]=head[h]; head[h]=p;
      // walk from p
      int c=prev[p]; int hops=0,tri=0; int first=-1;
      while(c>=0 && p-c<32506 && hops<128){ hops++; if(d[c]==d[p]&&d[c+1]==d[p+1]){tri++; if(first<0) first=hops;} c=prev[c]; }
      tot_h+=hops; tot_t+=tri; npos++; if(first>0){tot_first+=first; nfirst++;} else tot_first+=hops;
    }
  }
  printf("positions %ld  hash hops/pos %.2f  same-trigram hops/pos %.2f (%.1f%%)  hops until first same-trigram (or end) %.2f\n",npos,(double)tot_h/npos,(double)tot_t/npos,100.0*tot_t/tot_h,(double)tot_first/npos);
}
