/* Test harness: calls Java_com_github_sparkbwa_BwaJni_bwa_1jni through a fake JNIEnv function
 * table (only the four slots the shim uses are populated), standing in for a JVM.
 * usage: jni_fake <libbwa.so> arg0 arg1 ...   -> prints "rc=<n> released=<n>" */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
typedef const void* const* JNIEnv;
static char** g_args; static int g_n; static int g_released;
static const char* GetStringUTFChars(JNIEnv* e, void* s, unsigned char* c) { (void)e; (void)c; return (const char*)s; }
static void ReleaseStringUTFChars(JNIEnv* e, void* s, const char* u) { (void)e; (void)s; (void)u; ++g_released; }
static int GetArrayLength(JNIEnv* e, void* a) { (void)e; (void)a; return g_n; }
/* the literal argument "@NULL@" stands for a null String element */
static void* GetObjectArrayElement(JNIEnv* e, void* a, int i) { (void)e; (void)a; return strcmp(g_args[i], "@NULL@") ? g_args[i] : 0; }
int main(int argc, char** argv) {
  void* h = dlopen(argv[1], RTLD_NOW);
  if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 99; }
  int (*fn)(JNIEnv*, void*, int, void*, void*) = (int (*)(JNIEnv*, void*, int, void*, void*))dlsym(h, "Java_com_github_sparkbwa_BwaJni_bwa_1jni");
  if (!fn) { fprintf(stderr, "symbol missing\n"); return 98; }
  const void* table[256];
  memset(table, 0, sizeof(table));
  table[169] = (const void*)GetStringUTFChars; table[170] = (const void*)ReleaseStringUTFChars;
  table[171] = (const void*)GetArrayLength; table[173] = (const void*)GetObjectArrayElement;
  const void* const* env = table;
  g_args = argv + 2; g_n = argc - 2;
  int lens[1] = {0};
  int rc = fn(&env, 0, g_n, (void*)g_args, lens);
  printf("rc=%d released=%d\n", rc, g_released);
  return 0;
}
