"""The front-end's in-place tokenizer + decimal fast path (eqtlbma_bf_main.cpp: Span / split_spans / fast_atof) returns the
very doubles atof returns (the reference parses every cell with utils::split + atof, data_loader.cpp:437-527, 878-1010)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HARNESS = r'''
int main(){ mt19937_64 rng(1); long bad=0,n=0; char buf[64];
 const char *fixed[] = {"0","-0","0.0","000.5","1.","+3.25","inf","-inf","nan","1e5",".5","-.25","0.000000000000000000001",
                        "123456789012345.5","2","1.000","0.001","1999.999","NA"};
 for(int it=0;it<2000000;++it){ int kind=rng()%6;
  if(kind==0) snprintf(buf,64,"%.3f",(rng()%2001)/1000.0);
  else if(kind==1) snprintf(buf,64,"%.*f",(int)(rng()%18),((double)(rng()%2000000)-1e6)/977.0);
  else if(kind==2) snprintf(buf,64,"%.*e",(int)(rng()%17),((double)(rng()%2000000)-1e6)/977.0);
  else if(kind==3) snprintf(buf,64,"%lld",(long long)(rng()%100000000000000000ULL));
  else if(kind==4) snprintf(buf,64,"%.17g",(double)(rng()%1000003)/7.0);
  else snprintf(buf,64,"%s",fixed[rng()%19]);
  Span sp{buf,strlen(buf)}; double a=fast_atof(sp), b=atof(buf); ++n;
  if(memcmp(&a,&b,8)!=0 && !(a!=a && b!=b)){ if(bad<10) printf("MISMATCH %s %.17g %.17g\n",buf,a,b); ++bad; } }
 vector<Span> t; string line = "  snp1\t0.5  1.25\t\tNA 2 ";
 split_spans(line, t);
 if (t.size()!=5 || t[0].str()!="snp1" || !t[3].eq("NA") || !is_na(t[3]) || t[4].str()!="2") { printf("split_spans wrong\n"); ++bad; }
 printf("%ld checked, %ld mismatches\n",n,bad); return bad!=0; }
'''


def test_fast_atof_equals_atof(tmp_path):
    src = open(os.path.join(ROOT, "eqtlbma_b200", "host", "eqtlbma_bf_main.cpp")).read()
    code = src[src.index("struct Span {"):src.index("struct GzReader {")]
    cpp = tmp_path / "t.cpp"
    cpp.write_text("#include <cstring>\n#include <cstdlib>\n#include <cstdio>\n#include <string>\n#include <vector>\n"
                   "#include <random>\nusing namespace std;\n" + code + HARNESS)
    exe = str(tmp_path / "t")
    subprocess.check_call(["g++", "-O2", "-o", exe, str(cpp)])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout
