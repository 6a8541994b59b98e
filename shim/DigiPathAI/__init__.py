"""Drop-in shim package: put this directory's parent (``shim/``) in front of the reference on ``sys.path`` -- or copy
``Segmentation.py`` over ``DigiPathAI/Segmentation.py`` in an installed reference -- and the viewer's
``from DigiPathAI.Segmentation import getSegmentation`` (DigiPathAI/main_server.py:155, README.md:77) resolves to the
B200 path.  Nothing else of the reference package is shadowed on purpose when the file-level route is used."""
